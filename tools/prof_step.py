"""Warm per-kernel durations of one C2 scan step (CUPTI through torch.profiler): where the time outside the main kernel goes."""
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yams_b200 as Y  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
d, k = 768, 10
assert Y.plugin_init() == 0
c = Y.Corpus(d, Y.F16, Y.COSINE, capacity_hint=n)
for r0 in range(0, n, 1_000_000):
    c.append_synthetic(42, r0, min(1_000_000, n - r0))
q = torch.empty((nq, d), dtype=torch.float32, device="cuda")
Y.synth_rows_device(43, 0, nq, d, q.data_ptr())
out = torch.empty(nq * k * 12, dtype=torch.uint8, device="cuda")
for _ in range(5):
    c.search_device(q.data_ptr(), nq, k, -1.0, out.data_ptr(), out.data_ptr() + nq * k * 8)
c.sync()
steps = 10
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        c.search_device(q.data_ptr(), nq, k, -1.0, out.data_ptr(), out.data_ptr() + nq * k * 8)
    c.sync()
agg = collections.OrderedDict()
each = collections.defaultdict(list)     # individual durations, in launch order (the two selections of a step differ)
first, last = None, None
for e in prof.events():
    if e.device_type.name != "CUDA":
        continue
    a = agg.setdefault(e.name[:70], [0, 0.0])
    a[0] += 1
    dur = e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
    a[1] += dur
    each[e.name[:70]].append((e.time_range.start, dur))
    t0 = e.time_range.start
    t1 = e.time_range.end
    first = t0 if first is None else min(first, t0)
    last = t1 if last is None else max(last, t1)
tot = sum(v[1] for v in agg.values())
print(f"{steps} steps: wall {(last - first) / steps / 1e3:.3f} ms/step, kernel time sum {tot / steps / 1e3:.3f} ms/step, gaps {(last - first - tot) / steps / 1e3:.3f} ms/step")
for name, (cnt, us) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{us / steps:10.1f} us/step  x{cnt / steps:.0f}  {name}")
    if cnt / steps > 1:
        seq = [d for _, d in sorted(each[name])][:int(2 * cnt / steps)]
        print("            in launch order (us):", " ".join(f"{d:.0f}" for d in seq))
