"""Q sweep of the C2 scan (10M x 768 fp16): which roof binds and what fraction is achieved. Prints a markdown table."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = []
for q in [int(x) for x in (sys.argv[1:] or ["1", "8", "32", "64", "128", "256", "512", "1024"])]:
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "knn", "--queries", str(q), "--steps", "10",
                          "--warmup", "3", "--no-cpu-baseline"], capture_output=True, text=True)
    line = [l for l in out.stdout.splitlines() if l.startswith("{")]
    if not line:
        print("Q", q, "FAILED", out.stderr[-400:])
        continue
    d = json.loads(line[-1])
    r = d["roofline"]
    rows.append((q, d["value"], d["ms_per_step"], r["kernel_ms"], r["bound"], r["frac"], r["hbm_view"]["achieved"], r["hbm_view"]["frac"],
                 r["tensor_view"]["achieved"], r["tensor_view"]["frac"], d["e2e"]["value"]))
print("| Q | queries/s | ms/step | scan kernel ms | binding roof | frac of binding | HBM GB/s (frac) | tensor TF/s (frac) | e2e queries/s |")
print("|---:|---:|---:|---:|---|---:|---:|---:|---:|")
for r in rows:
    print(f"| {r[0]} | {r[1]:.0f} | {r[2]:.3f} | {r[3]:.3f} | {r[4]} | {r[5]:.3f} | {r[6]:.0f} ({r[7]:.3f}) | {r[8]:.0f} ({r[9]:.3f}) | {r[10]:.0f} |")
