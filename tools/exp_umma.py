"""Experiment driver for the tcgen05 stage-1 kernel (knn_umma.cu): one corpus, many kernel variants selected through the
YAMS_B200_UMMA_* environment knobs, the device time of the full-corpus filtered launch printed per variant.

  python tools/exp_umma.py [rows] [queries]
"""
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yams_b200 as Y  # noqa: E402

KNOBS = ["CTAS", "MODE", "NOPF", "PEER_ARRIVE", "NTILE", "STAGES", "PROF", "EPI_RELAXED"]


class Clocks:
    """SM clock sampled through NVML while a variant runs (power-capped parts: the clock IS the result)."""

    def __init__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(0)
        except Exception:  # noqa: BLE001
            self.nv = None
        self.samples, self.stop = [], False

    def _loop(self):
        while not self.stop:
            try:
                self.samples.append((self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM),
                                     self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.01)

    def __enter__(self):
        self.samples, self.stop = [], False
        if self.nv:
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        if self.nv:
            self.t.join()

    def summary(self):
        if not self.samples:
            return "clk n/a"
        tail = self.samples[len(self.samples) // 2:]
        return f"sm {np.median([s[0] for s in tail]):.0f} MHz {np.median([s[1] for s in tail]):.0f} W"


CLK = Clocks()


def run(c, q, k, reps=24, **kn):
    for name in KNOBS:
        os.environ.pop("YAMS_B200_UMMA_" + name, None)
    for name, v in kn.items():
        os.environ["YAMS_B200_UMMA_" + name.upper()] = str(v)
    ms = []
    with CLK:
        for _ in range(reps):
            res = c.search(q, k, threshold=-1.0)
            ms.append(c.last_timings()["scan_kernel_ms"])
    run.last = res
    return float(np.median(ms[reps // 2:])), [min(ms), max(ms), CLK.summary()]


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    d, k = 768, 10
    assert Y.plugin_init() == 0, Y.health()
    c = Y.Corpus(d, Y.F16, Y.COSINE, capacity_hint=rows)
    for r0 in range(0, rows, 1_000_000):
        c.append_synthetic(42, r0, min(1_000_000, rows - r0))
    rng = np.random.default_rng(7)
    q = rng.uniform(-1, 1, size=(nq, d)).astype(np.float32)
    flops = 2.0 * nq * rows * d
    variants = [
        dict(ctas=1),
        dict(ctas=2),
        dict(ctas=2, epi_relaxed=0),
        dict(ctas=2, nopf=1),
        dict(ctas=1, mode=1, prof=1),   # the experiment modes live in the diagnostic (PROF) instantiation
        dict(ctas=2, mode=1, prof=1),
        dict(ctas=1, mode=2, prof=1),
        dict(ctas=2, mode=2, prof=1),
        dict(ctas=2, stages=5),
        dict(ctas=2, stages=4),
        dict(ctas=2, stages=3),
    ]
    extra = os.environ.get("EXP_UMMA_VARIANTS")
    if extra:
        variants = [dict(kv.split("=") for kv in v.split(",")) for v in extra.split(";")]
    print(f"rows {rows} queries {nq} dim {d}: algorithmic {flops / 1e12:.2f} TFLOP per launch")
    ref = None
    for v in variants:
        try:
            best, info = run(c, q, k, reps=6 if v.get('mode') else 24, **v)
            same = ""
            if not v.get("mode"):
                if ref is None:
                    ref = run.last
                else:
                    same = " ids==ref %s scores==ref %s" % (np.array_equal(ref[0], run.last[0]), np.array_equal(ref[1], run.last[1]))
            print(f"{v!s:48s} kernel {best:8.3f} ms (median of the 2nd half) {flops / best / 1e9:8.1f} TF/s   min/max {info[0]:.3f}/{info[1]:.3f} {info[2]}{same}",
                  flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"{v!s:60s} FAILED {e!r}", flush=True)
    for v in (dict(ctas=1, prof=1), dict(ctas=2, prof=1)):
        print("profile", v, flush=True)
        sys.stderr.flush()
        run(c, q, k, reps=3, **v)
    c.close()


if __name__ == "__main__":
    main()
