#!/usr/bin/env python
"""Summarise Nsight Compute artefacts brought back in gpurun_out/ into small tracked files under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches_X.csv  > profiles/rN_launches_X.md
  python tools/ncu_summary.py full     gpurun_out/prof_X.ncu-rep   > profiles/rN_full_X.md
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/smem % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % active"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe % active"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe % active"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots % active"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__cycles_elapsed.max", "SM cycles"),
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = row["Metric Unit"]
        v *= {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# ncu launch list: `{path}`\n")
    print("`ncu --metrics gpu__time_duration.sum --clock-control none` — per-launch times are cold-cache and serialised: "
          "compare SHARES, not absolutes.\n")
    print("| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| `{k}` | {c} | {t/1e6:.3f} | {t/c/1e3:.1f} | {100*t/tot:.1f}% |")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full: `{path}`\n")
    for row in rows[2:]:
        name = row[hdr.index("Kernel Name")]
        print(f"## `{name[:110]}`\n")
        print("| metric | value |\n|---|---:|")
        for k, label in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"| {label} (`{k}`) | {row[i]} {units[i]} |")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
