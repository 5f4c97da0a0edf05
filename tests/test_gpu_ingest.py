"""GPU parity tests of the ingest path: every call goes through the C ABI of libyams_b200.so and is
compared bit-for-bit with the CPU oracle / the golden vectors generated from the reference."""
import hashlib

import numpy as np
import pytest

from tests.streams import cfg_from_dict, make_stream

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Y():
    import yams_b200
    assert yams_b200.device_count() > 0, "these tests need a B200"
    assert yams_b200.plugin_init() == 0, yams_b200.health()
    return yams_b200


def ycfg(Y, ocfg):
    return Y.CdcConfig(ocfg.window_size, ocfg.min_chunk, ocfg.max_chunk, ocfg.polynomial, ocfg.mask, ocfg.variant, 0)


# ---- SHA-256 ------------------------------------------------------------------------------------
def test_sha256_reference_kats(Y):
    # /root/reference/tests/unit/crypto/crypto_test.cpp:92-99
    base = np.frombuffer(b"abcHello World", dtype=np.uint8)
    d = Y.sha256_batch(base, [0, 0, 3], [0, 3, 11])
    assert bytes(d[0]).hex() == "e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855"
    assert bytes(d[1]).hex() == "ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad"
    assert bytes(d[2]).hex() == "a591a6d40bf420404a011733cfb7b190d62c65bf0bcda32b57b277d9ad9f146e"


def test_sha256_golden_sizes(Y, oracle, golden):
    for g in golden["sha256"]:
        data = oracle.gen_bytes(g["seed"], 0, max(g["size"], 1))[:g["size"]]
        d = Y.sha256_batch(data, [0], [g["size"]])
        assert bytes(d[0]).hex() == g["hex"], g["size"]


def test_sha256_every_length_and_alignment(Y, oracle):
    """All padding cases (len mod 64 in 0..63, around 55/56/64) at every byte alignment mod 16."""
    O = oracle
    base = O.gen_bytes(31337, 0, 1 << 16)
    offs, sizes = [], []
    for length in list(range(0, 200)) + [255, 256, 257, 1000, 4095, 4096, 4097]:
        for a in range(0, 16):
            offs.append(a + 17 * (length % 13))
            sizes.append(length)
    got = Y.sha256_batch(base, offs, sizes)
    want = O.sha256_batch(base, np.array(offs), np.array(sizes))
    assert np.array_equal(got, want)


def test_sha256_ragged_batch_many_chunks(Y, oracle):
    """More chunks than lanes in a wave, wildly different lengths -> work-queue refill paths."""
    O = oracle
    rng = np.random.default_rng(4)
    base = O.gen_bytes(8, 0, 24 << 20)
    n = 6000
    sizes = np.concatenate([rng.integers(0, 300, n // 2), rng.integers(1000, 70000, n // 2 - 4),
                            np.array([1 << 20, (1 << 20) + 1, 0, 64])]).astype(np.uint64)
    rng.shuffle(sizes)
    offs = rng.integers(0, base.size - (1 << 21), sizes.size).astype(np.uint64)
    got = Y.sha256_batch(base, offs, sizes)
    want = O.sha256_batch(base, offs, sizes)
    assert np.array_equal(got, want)
    assert bytes(got[0]) == hashlib.sha256(base[int(offs[0]):int(offs[0] + sizes[0])].tobytes()).digest()


# ---- CDC + SHA-256 --------------------------------------------------------------------------------
def test_chunk_and_hash_golden(Y, oracle, golden):
    for g in golden["cdc"]:
        data = make_stream(g["stream"], oracle)
        cfg = ycfg(Y, cfg_from_dict(oracle, g["cfg"]))
        ch = Y.chunk_and_hash(data, cfg)
        assert [int(x) for x in ch["offset"]] == g["offsets"], (g["stream"], g["config"])
        assert [int(x) for x in ch["size"]] == g["sizes"], (g["stream"], g["config"])
        assert [bytes(d).hex() for d in ch["digest"][:16]] == g["digests_head"]
        assert hashlib.sha256(np.ascontiguousarray(ch["digest"]).tobytes()).hexdigest() == g["digest_of_digests"]


def test_chunk_random_configs_vs_oracle(Y, oracle):
    O = oracle
    rng = np.random.default_rng(77)
    for trial in range(40):
        n = int(rng.integers(0, 400000))
        data = O.gen_bytes(int(rng.integers(1, 1 << 30)), 0, n)
        if trial % 3 == 0 and n:
            data[rng.integers(0, n, size=max(1, n // 40))] = 0xC5
        if trial % 7 == 0:
            data[:] = 0x42
        minc, maxc = int(rng.integers(0, 3000)), int(rng.integers(0, 9000))
        variant = trial % 2
        if variant == 1 and minc == 0 and maxc == 0:
            maxc = 1
        ocfg = O.default_config(variant=variant, min_chunk=minc, max_chunk=maxc, window_size=int(rng.integers(0, 49)),
                                mask=int(rng.choice([0x0, 0x3F, 0xFF, 0x1FF, 0x1FFF, 0x303, 0x10001, 0x8000000000000001])))
        if ocfg.mask == 0 and n > 50000:
            data = data[:50000]
        want = O.cdc_chunk(data, ocfg)
        got = Y.chunk_and_hash(data, ycfg(Y, ocfg))
        assert np.array_equal(got["offset"], want[0]) and np.array_equal(got["size"], want[1]), (trial, minc, maxc, variant)
        assert np.array_equal(got["digest"], want[2]), trial


def test_boundaries_only_and_invalid_config(Y, oracle):
    data = oracle.gen_bytes(3, 0, 300000)
    ocfg = oracle.default_config(min_chunk=2048, max_chunk=65536)
    b = Y.chunk_boundaries(data, ycfg(Y, ocfg))
    want = oracle.cdc_chunk(data, ocfg, hash=False)
    assert np.array_equal(b["offset"], want[0]) and np.array_equal(b["size"], want[1])
    assert not b["digest"].any()
    with pytest.raises(Y.YamsB200Error) as e:
        Y.chunk_and_hash(data, Y.default_config(window_size=49))
    assert e.value.status == 1
    with pytest.raises(Y.YamsB200Error):
        Y.chunk_and_hash(data, Y.default_config(variant=Y.RABIN, min_chunk=0, max_chunk=0))


def test_stream_fragmentation_is_invisible(Y, oracle):
    # /root/reference/tests/unit/chunking/chunking_test.cpp:481-523
    O = oracle
    data = O.gen_bytes(12345, 0, 3 << 20)
    ocfg = O.default_config(min_chunk=4096, max_chunk=65536)
    want = O.cdc_chunk(data, ocfg)
    rng = np.random.default_rng(9)
    for seg in (1 << 20, 65536, 4097, 1000):
        with Y.IngestSession(ycfg(Y, ocfg)) as s:
            parts = []
            pos = 0
            while pos < data.size:
                ln = seg if seg > 2000 else int(rng.integers(1, 3 * seg))
                parts.append(s.feed(data[pos:pos + ln]))
                pos += ln
            parts.append(s.finish())
        got = np.concatenate(parts)
        assert np.array_equal(got["offset"], want[0]) and np.array_equal(got["size"], want[1]), seg
        assert np.array_equal(got["digest"], want[2]), seg
    # empty stream, and finish without feed
    with Y.IngestSession() as s:
        assert len(s.feed(b"")) == 0 and len(s.finish()) == 0


def test_device_resident_matches_host_path_and_oracle(Y, oracle):
    import torch
    O = oracle
    n = (64 << 20) + 12345
    t = torch.empty(n, dtype=torch.uint8, device="cuda")
    Y.synth_bytes_device(12345, 0, n, t.data_ptr())
    host = t.cpu().numpy()
    assert np.array_equal(host[:1 << 20], O.gen_bytes(12345, 0, 1 << 20))
    assert np.array_equal(host[-4097:], O.gen_bytes(12345, n - 4097, 4097))
    for variant in (Y.STREAMING, Y.RABIN):
        got = Y.chunk_and_hash_device(t.data_ptr(), n, Y.default_config(variant=variant))
        want = O.cdc_chunk(host, O.default_config(variant=variant))
        assert np.array_equal(got["offset"], want[0]) and np.array_equal(got["size"], want[1])
        assert np.array_equal(got["digest"], want[2])
    # unaligned device sub-ranges (misaligned stream head), incl. inputs shorter than one 16-byte unit
    for off, ln in [(7, 5_000_001), (1, 100_000), (15, 33), (9, 5), (3, 16), (0, 1), (13, 70_000)]:
        for kw in (dict(min_chunk=1024, max_chunk=8192), dict(min_chunk=1, max_chunk=64, mask=0x3)):
            got = Y.chunk_and_hash_device(t.data_ptr() + off, ln, Y.default_config(**kw))
            want = O.cdc_chunk(host[off:off + ln], O.default_config(**kw))
            assert np.array_equal(got["offset"], want[0]) and np.array_equal(got["size"], want[1]), (off, ln, kw)
            assert np.array_equal(got["digest"], want[2]), (off, ln, kw)
    # a 0xC5 in the very first bytes of a misaligned stream (candidate inside the unaligned head)
    t2 = t[:200_000].clone()
    t2[3:12] = 0xC5
    h2 = t2.cpu().numpy()
    for off in (3, 5, 11):
        got = Y.chunk_and_hash_device(t2.data_ptr() + off, 150_000, Y.default_config(min_chunk=1, max_chunk=4096, mask=0xFF))
        want = O.cdc_chunk(h2[off:off + 150_000], O.default_config(min_chunk=1, max_chunk=4096, mask=0xFF))
        assert np.array_equal(got["offset"], want[0]) and np.array_equal(got["digest"], want[2]), off


def test_full_size_properties(Y, oracle):
    """BASELINE-size behaviour through size-independent properties: 6 GiB of the C3 stream resident in
    HBM (32-bit offsets overflow; segment boundaries are forced below): coverage, sequential offsets, size bounds,
    spot-checked digests, and equality of an interior window with the oracle re-run from a known cut."""
    import torch
    O = oracle
    n = 6 << 30
    t = torch.empty(n, dtype=torch.uint8, device="cuda")
    Y.synth_bytes_device(12345, 0, n, t.data_ptr())
    ch = Y.chunk_and_hash_device(t.data_ptr(), n, Y.default_config())
    offs, sizes = ch["offset"], ch["size"]
    assert offs[0] == 0 and int(offs[-1] + sizes[-1]) == n
    assert np.array_equal(offs[1:], (offs + sizes)[:-1])
    assert sizes[:-1].min() >= 16384 and sizes.max() <= 1 << 20
    assert 0.9 < (n / len(ch)) / 24576.0 < 1.1          # mean chunk ~ 16 KiB + 8 KiB
    rng = np.random.default_rng(1)
    for i in list(rng.integers(0, len(ch), 24)) + [0, len(ch) - 1]:
        o, s = int(offs[i]), int(sizes[i])
        seg = t[o:o + s].cpu().numpy().tobytes()
        assert hashlib.sha256(seg).digest() == bytes(ch["digest"][i]), i
    # chunking is a pure function of the stream prefix only through the rolling window: restarting the
    # oracle at a cut (with 48+8 bytes of history to seed its window) must reproduce the following cuts.
    j = int(np.searchsorted(offs, (4 << 30) + 12345))
    start = int(offs[j])
    span = 32 << 20
    hist = 64
    window = t[start - hist:start + span].cpu().numpy()
    want = O.cdc_candidates(window, O.default_config())  # candidates are position-local
    got_cuts = offs[j + 1:][offs[j + 1:] <= start + span] - 1
    cand_abs = want.astype(np.int64) + (start - hist)
    assert np.isin(got_cuts[sizes[j:j + len(got_cuts)] < (1 << 20)], cand_abs).all()
    # segmenting is invisible: the default (one 16 GiB segment here) and 1 / 2.5 GiB segments (the open chunk and the look-behind
    # cross every segment boundary) give the same table, digests included
    import os
    for mib in ("1024", "2560"):
        os.environ["YAMS_B200_SEGMENT_MIB"] = mib
        try:
            ch2 = Y.chunk_and_hash_device(t.data_ptr(), n, Y.default_config())
        finally:
            del os.environ["YAMS_B200_SEGMENT_MIB"]
        assert np.array_equal(ch2["offset"], offs) and np.array_equal(ch2["size"], sizes) and np.array_equal(ch2["digest"], ch["digest"]), mib


def test_dedup_stats_matches_reference_accounting(Y, oracle):
    """calculateDeduplication (/root/reference/src/chunking/rabin_chunker.cpp:224-239): unique-hash accounting."""
    O = oracle
    block = O.gen_bytes(21, 0, 1 << 20)
    data = np.concatenate([block, block, O.gen_bytes(22, 0, 300_000), block[:500_000], np.zeros(0, dtype=np.uint8)])
    ch = Y.chunk_and_hash(data, Y.default_config(min_chunk=2048, max_chunk=32768))
    got = Y.dedup_stats(ch)
    seen, usz = set(), 0
    for c in ch:
        key = bytes(c["digest"])
        if key not in seen:
            seen.add(key)
            usz += int(c["size"])
    assert got["chunkCount"] == len(ch) and got["totalSize"] == data.size
    assert got["uniqueChunks"] == len(seen) and got["uniqueSize"] == usz
    assert got["uniqueChunks"] < got["chunkCount"]          # the repeated megabyte dedups
    assert Y.dedup_stats(ch[:0])["chunkCount"] == 0


def test_digest_set_matches_sequential_exists_store_loop(Y, oracle):
    """content_store_impl.cpp:245-288: per chunk `exists(hash)` then store -- replayed on a Python set."""
    rng = np.random.default_rng(9)
    pool = rng.integers(0, 256, size=(5000, 32), dtype=np.uint8)
    pool[1, :16] = pool[0, :16]                       # same hash-slot key, different digest
    s = Y.DigestSet()
    model = set()
    assert len(s.contains(pool[:0])) == 0
    for rnd, n in enumerate((1, 7, 1000, 20000, 3, 60000)):
        batch = pool[rng.integers(0, min(len(pool), 50 + 2000 * rnd), size=n)]
        before = s.contains(batch)
        assert list(before) == [bytes(d) in model for d in batch]
        existed, new = s.insert(batch)
        want = []
        for d in batch:
            b = bytes(d)
            want.append(b in model)
            model.add(b)
        assert list(existed) == want
        assert new == want.count(False) and len(s) == len(model)
        assert np.all(s.contains(batch) == 1)
    assert np.all(s.contains(rng.integers(0, 256, size=(1000, 32), dtype=np.uint8)) == 0)
    s.close()
    # chunk tables go in directly (stride = sizeof(yams_chunk_desc)); agrees with calculateDeduplication
    block = oracle.gen_bytes(5, 0, 1 << 20)
    data = np.concatenate([block, block, oracle.gen_bytes(6, 0, 1 << 19), block])
    ch = Y.chunk_and_hash(data, Y.default_config(min_chunk=2048, max_chunk=32768, mask=0x3FF))
    s = Y.DigestSet(capacity_hint=16)
    existed, new = s.insert(ch)
    st = Y.dedup_stats(ch)
    assert new == st["uniqueChunks"] == len(s)
    assert int(ch["size"][existed == 0].sum()) == st["uniqueSize"]
    existed2, new2 = s.insert(ch)
    assert new2 == 0 and np.all(existed2 == 1)
    s.close()


def _batch_files(O):
    rng = np.random.default_rng(23)
    sizes = [0, 1, 47, 48, 49, 100, 4095, 4096, 16383, 16384, 16385, 65536, (1 << 20) + 3, 0, 5 << 20, 333_333]
    files = []
    for i, n in enumerate(sizes):
        kind = i % 4
        if kind == 0:
            files.append(O.gen_bytes(100 + i, 0, n))
        elif kind == 1:
            files.append(np.full(n, 0x42, dtype=np.uint8))                                   # no candidates at all
        elif kind == 2:
            files.append(((np.arange(n, dtype=np.uint64) * 1315423911 + 0x9E3779B9) & 0xFF).astype(np.uint8))   # reference pattern
        else:
            files.append(rng.integers(0, 256, size=n, dtype=np.uint8))
    files.append(files[12].copy())                                                           # a repeated file
    return files


def _check_batch(Y, O, files, cfg):
    got = Y.chunk_and_hash_batch(files, cfg)
    assert len(got) == len(files)
    ocfg = O.CdcConfig(cfg.window_size, cfg.min_chunk, cfg.max_chunk, cfg.polynomial, cfg.mask, cfg.variant)
    for i, f in enumerate(files):
        off, size, dig = O.cdc_chunk(f, ocfg)
        g = got[i]
        assert len(g) == len(off), (i, len(f), len(g), len(off))
        assert np.array_equal(g["offset"], off) and np.array_equal(g["size"], size), (i, len(f))
        assert np.array_equal(g["digest"], dig), (i, len(f))


def test_chunk_and_hash_batch_equals_per_file_reference(Y, oracle):
    """Every file of a batch is an independent stream: identical to chunking + hashing the files one by one
    (rabin_chunker.cpp:120-152 / streaming_chunker.h:146-204 per file)."""
    O = oracle
    files = _batch_files(O)
    for variant in (Y.STREAMING, Y.RABIN):
        _check_batch(Y, O, files, Y.default_config(variant))
        _check_batch(Y, O, files, Y.default_config(variant, min_chunk=64, max_chunk=1024, mask=0x3F))
        _check_batch(Y, O, files, Y.default_config(variant, min_chunk=4096, max_chunk=65536))
    _check_batch(Y, O, files, Y.default_config(Y.STREAMING, min_chunk=2048, max_chunk=512))      # min >= max: forced cuts only
    _check_batch(Y, O, files, Y.default_config(Y.STREAMING, min_chunk=0, max_chunk=4096, mask=0xFF))
    small = [f[:3000] for f in files]
    _check_batch(Y, O, small, Y.default_config(Y.STREAMING, min_chunk=0, max_chunk=700, mask=0x0))          # every position is a candidate
    _check_batch(Y, O, small, Y.default_config(Y.RABIN, min_chunk=5, max_chunk=900, mask=0x0, window_size=7))
    assert Y.chunk_and_hash_batch([], Y.default_config()) == []
    one = Y.chunk_and_hash_batch([files[12]], Y.default_config())
    assert np.array_equal(one[0], Y.chunk_and_hash(files[12], Y.default_config()))


def test_chunk_and_hash_batch_group_splitting(oracle):
    """A 1 MiB staging budget forces several groups and sends the larger files through the streaming path."""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import numpy as np, yams_b200 as Y\n"
        "from oracle import oracle as O\n"
        "from tests.test_gpu_ingest import _batch_files, _check_batch\n"
        "assert Y.plugin_init() == 0\n"
        "files = _batch_files(O) * 2\n"
        "_check_batch(Y, O, files, Y.default_config())\n"
        "_check_batch(Y, O, files, Y.default_config(Y.RABIN, min_chunk=256, max_chunk=8192, mask=0xFF))\n"
        "print('BATCH OK')\n" % root)
    env = dict(os.environ, YAMS_B200_BATCH_MIB="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert "BATCH OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_sha256_many_separate_messages(Y, oracle):
    """IContentHasher::hash over separately held spans / ChunkValidator::validateChunks (chunk_validator.cpp:173-213)."""
    O = oracle
    rng = np.random.default_rng(31)
    sizes = [0, 1, 55, 56, 63, 64, 65, 119, 120, 1000, 0, 65536, (1 << 20) + 7, 5 << 20, 3]
    msgs = [rng.integers(0, 256, size=n, dtype=np.uint8) for n in sizes]
    got = Y.sha256_many(msgs)
    for i, m in enumerate(msgs):
        assert bytes(got[i]) == O.sha256(m), (i, sizes[i])
    assert Y.sha256_many([]).shape == (0, 32)
    assert bytes(Y.sha256_many([b"abc"])[0]).hex() == "ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad"


def test_sha256_longest_first_order_covers_every_bucket(Y, oracle):
    """Tables of >= 4096 chunks are hashed longest first (32 size buckets, sha_order_*_kernel): sizes spanning every bucket, many
    equal sizes, empty messages -- every digest must still land in its own slot."""
    rng = np.random.default_rng(41)
    n = 6000
    sizes = np.concatenate([rng.integers(0, 200, 1500), rng.integers(200, 20000, 3000), rng.integers(20000, 90000, 1400),
                            np.full(96, 4096), np.array([0, 0, 300000, 1 << 20])]).astype(np.int64)
    rng.shuffle(sizes)
    assert sizes.size == n
    pool = rng.integers(0, 256, size=int(sizes.max()) + n, dtype=np.uint8)
    msgs = [pool[i:i + int(s)] for i, s in enumerate(sizes)]          # overlapping windows of one pool: distinct contents, little memory
    got = Y.sha256_many(msgs)
    for i in list(range(0, n, 7)) + [int(np.argmax(sizes)), int(np.argmin(sizes))]:
        assert bytes(got[i]) == hashlib.sha256(msgs[i].tobytes()).digest(), (i, int(sizes[i]))
    # the chunker's own table: a device-resident stream with > 4096 chunks, every digest against hashlib
    import torch
    ln = 96 << 20
    t = torch.empty(ln, dtype=torch.uint8, device="cuda")
    Y.synth_bytes_device(77, 0, ln, t.data_ptr())
    got = Y.chunk_and_hash_device(t.data_ptr(), ln, Y.default_config(min_chunk=2048, max_chunk=65536, mask=0x7FF))
    host = t.cpu().numpy()
    assert len(got) > 4096
    for i in range(0, len(got), 5):
        o, s = int(got["offset"][i]), int(got["size"][i])
        assert bytes(got["digest"][i]) == hashlib.sha256(host[o:o + s].tobytes()).digest(), i
