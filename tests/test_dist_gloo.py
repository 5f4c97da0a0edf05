"""world_size-2 gloo test (CPU) of the N>1 host logic: range sharding + all-gather layout + merge order.
Each rank scores its shard with the oracle, pads to the [Q][k] partial layout the device kernels produce,
all-gathers, and the merged result must equal the oracle's answer over the whole corpus."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _merge_like_device(all_r, all_s, k):
    """Re-statement of merge_partials_kernel's order: (score desc, rowid asc), rowid -1 = padding."""
    R, Q, _ = all_r.shape
    out_r = np.full((Q, k), -1, dtype=np.int64)
    out_s = np.full((Q, k), -np.inf, dtype=np.float32)
    for q in range(Q):
        r = all_r[:, q, :].reshape(-1)
        s = all_s[:, q, :].reshape(-1)
        keep = r >= 0
        r, s = r[keep], s[keep]
        order = np.lexsort((r, -s.astype(np.float64)))[:k]
        out_r[q, :len(order)] = r[order]
        out_s[q, :len(order)] = s[order]
    return out_r, out_s


def _candidate_lists(n, nq):
    rng = np.random.default_rng(99)
    lists = [np.sort(rng.choice(n, size=sz, replace=False)).astype(np.int64) for sz in (1, 5, 40, 300, 1500, n)][:nq]
    lists[1] = np.zeros(0, dtype=np.int64)
    return lists


def _worker(rank, world, port, n, d, nq, k, ret, with_candidates=False):
    import torch
    import torch.distributed as dist
    from oracle import oracle as O
    from yams_b200.dist import allgather_partials, shard_rows
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, cnt = shard_rows(n, rank, world)
    rows = O.f16_from_float(O.gen_rows_f32(42, first, cnt, d)).reshape(cnt, d)
    queries = O.gen_rows_f32(43, 0, nq, d)
    part_r = np.full((nq, k), -1, dtype=np.int64)
    part_s = np.full((nq, k), -np.inf, dtype=np.float32)
    mine = None
    if with_candidates:
        from yams_b200.dist import split_allowed
        mine = split_allowed(_candidate_lists(n, nq), first, first + cnt)
        assert all(len(a) == 0 or (a[0] >= first and a[-1] < first + cnt) for a in mine)
    for q in range(nq):
        if mine is not None and len(mine[q]) == 0:
            continue                                   # nothing of this query's candidate set lives in this shard
        rc, r, s = O.exact_scan_cosine(rows, queries[q], k, threshold=-1.0, rowids=np.arange(first, first + cnt),
                                       allowed=mine[q] if mine is not None else None)
        part_r[q, :len(r)] = r
        part_s[q, :len(s)] = s
    all_r, all_s = allgather_partials(torch.from_numpy(part_r), torch.from_numpy(part_s))
    assert tuple(all_r.shape) == (world, nq, k)
    mr, ms = _merge_like_device(all_r.numpy(), all_s.numpy(), k)
    if rank == 0:
        ret["r"], ret["s"] = mr, ms
    dist.barrier()
    dist.destroy_process_group()


def test_shard_rows_partition():
    from yams_b200.dist import shard_rows
    for n in (0, 1, 7, 100, 10_000_001):
        for w in (1, 2, 3, 8):
            spans = [shard_rows(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_two_rank_gather_merge_equals_global(oracle):
    import torch.multiprocessing as mp
    O = oracle
    n, d, nq, k, world = 3001, 32, 6, 10, 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, d, nq, k, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    rows = O.f16_from_float(O.gen_rows_f32(42, 0, n, d)).reshape(n, d)
    queries = O.gen_rows_f32(43, 0, nq, d)
    for q in range(nq):
        rc, r, s = O.exact_scan_cosine(rows, queries[q], k, threshold=-1.0)
        assert list(ret["r"][q]) == list(r)
        assert np.array_equal(ret["s"][q], s)


def test_two_rank_candidate_sets_equal_global(oracle):
    """Config C5 on a sharded corpus: allowed-rowid lists split by shard range, partial top-k gathered and merged."""
    import torch.multiprocessing as mp
    O = oracle
    n, d, nq, k, world = 3001, 32, 6, 10, 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, d, nq, k, ret, True)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    rows = O.f16_from_float(O.gen_rows_f32(42, 0, n, d)).reshape(n, d)
    queries = O.gen_rows_f32(43, 0, nq, d)
    lists = _candidate_lists(n, nq)
    for q in range(nq):
        if len(lists[q]) == 0:
            assert np.all(ret["r"][q] == -1)
            continue
        rc, r, s = O.exact_scan_cosine(rows, queries[q], k, threshold=-1.0, allowed=lists[q])
        assert list(ret["r"][q][:len(r)]) == list(r) and np.all(ret["r"][q][len(r):] == -1)
        assert np.array_equal(ret["s"][q][:len(s)], s)
