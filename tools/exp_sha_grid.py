"""SHA-256 throughput vs resident CTAs per SM (is there room to co-schedule the candidate scan?)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, torch
sys.path.insert(0, %r)
import yams_b200 as Y
assert Y.plugin_init() == 0
n = 16 << 30
buf = torch.empty(n, dtype=torch.uint8, device="cuda")
Y.synth_bytes_device(12345, 0, n, buf.data_ptr())
for _ in range(3):
    ch = Y.chunk_and_hash_device(buf.data_ptr(), n, Y.default_config())
t = Y.ingest_last_timings()
print("sha256 %%.2f ms = %%.0f GB/s ; scan %%.2f select %%.2f total %%.2f" %% (t["sha256_ms"], n / t["sha256_ms"] / 1e6, t["scan_ms"], t["select_ms"], t["total_ms"]))
''' % ROOT
for grid in ("", "4", "3", "2"):
    env = dict(os.environ)
    if grid:
        env["YAMS_B200_SHA_GRID"] = grid
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print("SHA_GRID=%s" % (grid or "default(4)"), out.stdout.strip() or out.stderr[-300:], flush=True)
