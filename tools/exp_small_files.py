"""`yams add` of many small files: per-call latency of chunk_and_hash (host buffer in, chunk table out) by file size, and
aggregate throughput when several host threads ingest independent files concurrently (the reference parallelises
ingest across files/workers the same way).  Prints a markdown table."""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import yams_b200 as Y
from oracle import oracle as O

assert Y.plugin_init() == 0
cfg = Y.default_config()
print("| file size | threads | calls | ms per call | files/s | aggregate GB/s |")
print("|---:|---:|---:|---:|---:|---:|")
for size in (64 << 10, 1 << 20, 16 << 20, 256 << 20):
    data = torch.from_numpy(O.gen_bytes(7, 0, size)).pin_memory().numpy()
    for nthreads in (1, 4, 16):
        calls = max(4, min(200, (1 << 30) // size // nthreads))
        for _ in range(3):
            Y.chunk_and_hash(data, cfg)

        def work():
            for _ in range(calls):
                Y.chunk_and_hash(data, cfg)

        ts = [threading.Thread(target=work) for _ in range(nthreads)]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        dt = time.perf_counter() - t0
        total = calls * nthreads
        print(f"| {size >> 10} KiB | {nthreads} | {total} | {dt / calls * 1e3:.3f} | {total / dt:.0f} | {total * size / dt / 1e9:.2f} |")
