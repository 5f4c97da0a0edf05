"""Generate tests/golden/*.json by EXECUTING THE REFERENCE (oracle/_ref, built from /root/reference
by oracle/Makefile).  Run in the build container only:  python tests/golden/make_golden.py

The fixtures pin oracle/yams_oracle.c and the CUDA path on machines where /root/reference does not
exist (the GPU box).  Inputs are regenerated from seeds; only outputs are stored.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import oracle as O  # noqa: E402


def pattern_bytes(n):
    # tests/unit/chunking/chunking_test.cpp:55-62
    i = np.arange(n, dtype=np.uint64)
    return ((i * np.uint64(1315423911) + np.uint64(0x9E3779B9)) & np.uint64(0xFF)).astype(np.uint8)


def streams():
    yield "splitmix_12345_8MiB", O.gen_bytes(12345, 0, 8 << 20)
    yield "pattern_4MiB", pattern_bytes(4 << 20)
    yield "const42_3MiB_plus7", np.full((3 << 20) + 7, 0x42, dtype=np.uint8)
    yield "splitmix_7_100000", O.gen_bytes(7, 0, 100000)
    yield "empty", np.zeros(0, dtype=np.uint8)
    yield "one_byte", np.array([0xC5], dtype=np.uint8)


CONFIGS = {
    "default_streaming": dict(variant=O.STREAMING),
    "default_rabin": dict(variant=O.RABIN),
    # tests/benchmarks/core_benchmarks.cpp:226-227 (min 4 KiB / max 64 KiB, RabinChunker)
    "bench_rabin_4k_64k": dict(variant=O.RABIN, min_chunk=4096, max_chunk=65536),
    "bench_streaming_4k_64k": dict(variant=O.STREAMING, min_chunk=4096, max_chunk=65536),
    # short window / wider mask / other polynomial exercise the general predicate
    "w16_mask0fff_min2k_max32k": dict(variant=O.STREAMING, window_size=16, mask=0x0FFF,
                                      min_chunk=2048, max_chunk=32768),
    "mask_3byte": dict(variant=O.RABIN, mask=0x10101, min_chunk=1024, max_chunk=16384),
    "poly_alt": dict(variant=O.STREAMING, polynomial=0xBFE6B8A5BF378D83, mask=0x7FF, min_chunk=512,
                     max_chunk=8192),
}


def main():
    assert O.ref_available(), "build oracle/_ref first (make -C oracle ref)"
    R = O.ref()
    out = {"generator": "tests/golden/make_golden.py", "reference": "trvon/yams@8ab82c1c", "cdc": []}
    for sname, data in streams():
        for cname, kw in CONFIGS.items():
            cfg = O.default_config(**kw)
            offs, sizes, dig = O.ref_chunk(data, cfg)
            out["cdc"].append({
                "stream": sname, "config": cname,
                "cfg": {k: int(getattr(cfg, k)) for k, _ in cfg._fields_},
                "offsets": [int(x) for x in offs], "sizes": [int(x) for x in sizes],
                # first 16 digests verbatim + SHA-256 over the concatenation of all of them
                "digests_head": [bytes(d).hex() for d in dig[:16]],
                "digest_of_digests": hashlib.sha256(dig.tobytes()).hexdigest(),
            })
    # SHA-256: reference one-shot over seeded buffers at the sizes its tests use
    # (tests/unit/crypto/crypto_test.cpp:172-187) plus padding-boundary sizes
    sha = []
    import ctypes as C
    for size in [0, 1, 17, 55, 56, 57, 63, 64, 65, 119, 120, 4096, 65537, 1 << 20]:
        data = O.gen_bytes(99, 0, size)
        buf = C.create_string_buffer(65)
        R.ref_sha256_hex(O._data_ptr(data), size, buf)
        sha.append({"seed": 99, "size": size, "hex": buf.value.decode()})
    out["sha256"] = sha
    # fp16 truncating conversion
    vals = np.array([0.0, -0.0, 1.0, -1.0, 0.1, 0.333333, 65504.0, 65520.0, 1e-8, 6e-8, 6.1e-5,
                     3.0e-5, 1e5, -1e5, 0.99999, 1.00097, 2.5e-7, np.inf, -np.inf], dtype=np.float32)
    out["f16"] = {"inputs_bits": [int(x) for x in vals.view(np.uint32)],
                  "half_bits": [int(R.ref_f16_from_float(float(v))) for v in vals]}
    # distances: AVX paths of the reference on seeded vectors
    dist = []
    for d in [3, 8, 16, 17, 128, 384, 768]:
        a = O.gen_rows_f32(1, 0, 1, d)[0] * 3.0
        b = O.gen_rows_f32(2, 0, 1, d)[0] * 0.5
        dist.append({"d": d, "l2": float(R.ref_l2_distance(O._p(a, O.f32p), O._p(b, O.f32p), d)),
                     "cosine": float(R.ref_cosine_distance(O._p(a, O.f32p), O._p(b, O.f32p), d))})
    out["distances"] = dist
    # C1 config: 1000x128 fp32, 16 queries, top-10 (sqlite-vec-cpp cosine + partial_sort)
    rows = O.gen_rows_f32(42, 0, 1000, 128)
    queries = O.gen_rows_f32(43, 0, 16, 128)
    c1 = []
    for q in queries:
        idx, dd = O.batch_top_k(q, rows, 10, O.METRIC_COSINE, use_ref=True)
        c1.append({"idx": [int(i) for i in idx], "dist": [float(x) for x in dd]})
    out["c1_cosine_top10"] = c1
    with open(os.path.join(HERE, "reference_golden.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", os.path.join(HERE, "reference_golden.json"),
          os.path.getsize(os.path.join(HERE, "reference_golden.json")), "bytes")


if __name__ == "__main__":
    main()
