"""SimeonPqAdc engine timing: filtered (sampled threshold) vs unfiltered (per-tile sorting network) ADC pass, 10 M x 768 fp16 rows."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, time, numpy as np
sys.path.insert(0, %r)
import yams_b200 as Y
assert Y.plugin_init() == 0
n, d, m, kc = int(sys.argv[1]), 768, 32, 256
c = Y.Corpus(d, Y.F16, Y.COSINE, capacity_hint=n)
for r0 in range(0, n, 1_000_000):
    c.append_synthetic(42, r0, min(1_000_000, n - r0))
rng = np.random.default_rng(5)
cb = (rng.normal(size=(m, kc, d // m)) / np.sqrt(d)).astype(np.float32).reshape(-1)
t0 = time.perf_counter(); pq = Y.PqIndex(c, m, kc, cb); build = time.perf_counter() - t0
out = []
for nq in (1, 2, 8, 64):
    q = rng.uniform(-1, 1, size=(nq, d)).astype(np.float32)
    for k, rr in ((10, 2), (100, 10)):
        res = pq.search(q, k, rerank_factor=rr, threshold=-1.0)
        t0 = time.perf_counter()
        for _ in range(5): res = pq.search(q, k, rerank_factor=rr, threshold=-1.0)
        ms = (time.perf_counter() - t0) / 5 * 1e3
        out.append("nq=%%d k=%%d x%%d: %%.3f ms/call (%%.0f GB/s of codes) sum=%%d" %% (nq, k, rr, ms, n * m * nq / ms / 1e6, int(res[0][:, :k].sum())))
print("build %%.2fs | " %% build + " | ".join(out))
''' % ROOT
n = sys.argv[1] if len(sys.argv) > 1 else "10000000"
for extra in ({}, {"YAMS_B200_PQ_UNFILTERED": "1"}):
    env = dict(os.environ)
    env.update(extra)
    out = subprocess.run([sys.executable, "-c", code, n], env=env, capture_output=True, text=True)
    print("unfiltered" if extra else "filtered  ", "|", out.stdout.strip() or out.stderr[-600:], flush=True)
