"""CPU check of the GPU chunker's integer logic (yams_b200/csrc/cdc_logic.h) through the test-only
host emulation tests/sim/cdc_sim.cpp, against the oracle.  No GPU needed; the product library is not
involved."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests.streams import cfg_from_dict, make_stream

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def sim():
    so = os.path.join(HERE, "sim", "libcdc_sim.so")
    src = os.path.join(HERE, "sim", "cdc_sim.cpp")
    hdr = os.path.join(HERE, "..", "yams_b200", "csrc", "cdc_logic.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src])
    L = C.CDLL(so)
    u8p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)
    L.sim_chunk.restype = C.c_size_t
    L.sim_chunk.argtypes = [u8p, C.c_size_t, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int,
                            C.c_size_t, u64p, u64p, C.c_size_t]
    L.sim_chunk_batch.restype = C.c_size_t
    L.sim_chunk_batch.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_size_t, C.c_uint64, C.c_uint64, C.c_uint64,
                                  C.c_uint64, C.c_uint64, C.c_int, u64p, u64p, C.c_size_t, u64p]
    L.sim_candidates.restype = C.c_size_t
    L.sim_candidates.argtypes = [u8p, C.c_size_t, C.c_uint64, C.c_uint64, C.c_uint64, u64p, C.c_size_t]
    return L


def run_sim(L, O, data, cfg, slice_=0):
    cap = data.size + 16
    offs = np.empty(cap, dtype=np.uint64)
    sizes = np.empty(cap, dtype=np.uint64)
    n = L.sim_chunk(O._data_ptr(data), data.size, cfg.window_size, cfg.min_chunk, cfg.max_chunk, cfg.polynomial,
                    cfg.mask, cfg.variant, slice_, O._p(offs, O.u64p), O._p(sizes, O.u64p), cap)
    assert n != 2**64 - 1
    return offs[:n], sizes[:n]


def test_sim_matches_golden(sim, oracle, golden):
    O = oracle
    for g in golden["cdc"]:
        if g["stream"] == "splitmix_12345_8MiB" and g["config"] not in ("default_streaming", "default_rabin"):
            continue
        data = make_stream(g["stream"], O)
        cfg = cfg_from_dict(O, g["cfg"])
        offs, sizes = run_sim(sim, O, data, cfg)
        assert [int(x) for x in offs] == g["offsets"], (g["stream"], g["config"])
        assert [int(x) for x in sizes] == g["sizes"], (g["stream"], g["config"])


def test_sim_random_configs_and_fragmentation(sim, oracle):
    O = oracle
    rng = np.random.default_rng(77)
    for trial in range(60):
        n = int(rng.integers(0, 120000))
        data = O.gen_bytes(int(rng.integers(1, 1 << 30)), 0, n)
        if trial % 3 == 0 and n:
            data[rng.integers(0, n, size=max(1, n // 40))] = 0xC5
        if trial % 7 == 0:
            data[:] = 0x42
        minc = int(rng.integers(0, 3000))
        maxc = int(rng.integers(0, 9000))
        variant = int(trial % 2)
        if variant == 1 and minc == 0 and maxc == 0:
            maxc = 1
        cfg = O.default_config(variant=variant, min_chunk=minc, max_chunk=maxc,
                               window_size=int(rng.integers(0, 49)),
                               mask=int(rng.choice([0x0, 0x3F, 0xFF, 0x1FF, 0x1FFF, 0x303, 0x10001, 0x8000000000000001])))
        want = O.cdc_chunk(data, cfg, hash=False)
        for slice_ in (0, 1 + int(rng.integers(0, 5000)), 65536):
            offs, sizes = run_sim(sim, O, data, cfg, slice_)
            assert np.array_equal(offs, want[0]) and np.array_equal(sizes, want[1]), (trial, slice_, minc, maxc, variant)


def test_sim_many_candidates_cross_blocks(sim, oracle):
    """> kNodeBlock candidates so the block exit / walk / mark path is exercised."""
    O = oracle
    data = O.gen_bytes(5, 0, 1 << 20)
    cfg = O.default_config(mask=0x3F, min_chunk=64, max_chunk=1024)   # ~1/64 density -> ~16k candidates
    want = O.cdc_chunk(data, cfg, hash=False)
    assert len(O.cdc_candidates(data, cfg)) > 4096
    for slice_ in (0, 100000):
        offs, sizes = run_sim(sim, O, data, cfg, slice_)
        assert np.array_equal(offs, want[0]) and np.array_equal(sizes, want[1])


def test_sim_batch_of_files_equals_per_file_oracle(sim, oracle):
    """chunk_and_hash_batch's selection logic (merged node table, dead links in the gaps, per-file tails) on the CPU."""
    O = oracle
    rng = np.random.default_rng(123)
    for trial in range(25):
        nfiles = int(rng.integers(1, 12))
        files = []
        for f in range(nfiles):
            n = int(rng.choice([0, 1, 47, 48, 49, 500, 4096, int(rng.integers(0, 60000))]))
            kind = int(rng.integers(0, 3))
            if kind == 0:
                files.append(O.gen_bytes(int(rng.integers(1, 1 << 30)), 0, n))
            elif kind == 1:
                files.append(np.full(n, int(rng.choice([0x00, 0x42, 0xC5])), dtype=np.uint8))
            else:
                d = O.gen_bytes(int(rng.integers(1, 1 << 30)), 0, n)
                if n:
                    d[rng.integers(0, n, size=max(1, n // 30))] = 0xC5
                files.append(d)
        minc, maxc = int(rng.integers(0, 3000)), int(rng.integers(0, 9000))
        variant = trial % 2
        if variant == 1 and minc == 0 and maxc == 0:
            maxc = 1
        cfg = O.default_config(variant=variant, min_chunk=minc, max_chunk=maxc, window_size=int(rng.integers(0, 49)),
                               mask=int(rng.choice([0x0, 0x3F, 0xFF, 0x1FFF, 0x303, 0x8000000000000001])))
        if cfg.mask == 0:
            files = [f[:3000] for f in files]
        total = sum(f.size for f in files)
        cap = total + 16 * nfiles + 16
        offs = np.empty(cap, dtype=np.uint64)
        sizes = np.empty(cap, dtype=np.uint64)
        first = np.zeros(nfiles + 1, dtype=np.uint64)
        ptrs = (C.c_void_p * nfiles)(*[f.ctypes.data if f.size else None for f in files])
        lens = (C.c_size_t * nfiles)(*[f.size for f in files])
        n = sim.sim_chunk_batch(ptrs, lens, nfiles, cfg.window_size, cfg.min_chunk, cfg.max_chunk, cfg.polynomial, cfg.mask,
                                cfg.variant, O._p(offs, O.u64p), O._p(sizes, O.u64p), cap, O._p(first, O.u64p))
        assert n < 2**63, (trial, n)
        assert int(first[nfiles]) == n
        for f in range(nfiles):
            want = O.cdc_chunk(files[f], cfg, hash=False)
            a, b = int(first[f]), int(first[f + 1])
            assert np.array_equal(offs[a:b], want[0]) and np.array_equal(sizes[a:b], want[1]), (trial, f, files[f].size, minc, maxc, variant)
