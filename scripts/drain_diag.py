"""What the tensor-core kernel's selection epilogue costs, with SM clock and board power sampled while it runs (dev tool).

gemm_debug 1 = nothing compared (TMEM loads only; results are wrong), 0 = the product.  20 batches back to back per
mode: the kernel runs at the board's power limit, so the sustained figures are what a serving loop sees.
(profiles/r02_drain_diag.jsonl additionally holds one-off modes of a diagnostic build: 16 = pre-test constants loaded,
nothing compared; 8 = pre-test computed, hits ignored; 12 = the same with the constants in registers.)"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlvectordb_b200 import DeviceShard  # noqa: E402
from oracle import synthetic  # noqa: E402

import threading
import time

import pynvml

pynvml.nvmlInit()
_dev = pynvml.nvmlDeviceGetHandleByIndex(0)


class Sampler:
    """SM clock / board power while a mode runs (NVML, every 5 ms)."""

    def __enter__(self):
        self.stop = False
        self.mhz, self.watts = [], []
        self.t = threading.Thread(target=self.run)
        self.t.start()
        return self

    def run(self):
        while not self.stop:
            self.mhz.append(pynvml.nvmlDeviceGetClockInfo(_dev, pynvml.NVML_CLOCK_SM))
            self.watts.append(pynvml.nvmlDeviceGetPowerUsage(_dev) / 1e3)
            time.sleep(0.005)

    def __exit__(self, *exc):
        self.stop = True
        self.t.join()


ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=4_000_000)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--nq", type=int, default=4096)
ap.add_argument("--k", type=int, default=100)
a = ap.parse_args()
for space in ("l2", "cosine"):
    s = DeviceShard(a.dim, space, capacity=a.rows)
    s.add_synthetic(42, 0, a.rows, space != "l2")
    Q = synthetic.queries(43, a.nq, a.dim)
    s.set_timing(True)
    s.set_tuning("gemm", 1)
    s.set_tuning("gemm_passes", 2)
    for wide in (3,):
        s.set_tuning("gemm_wide", wide)
        for passes, dbg in ((2, 0), (2, 1), (0, 0)):   # (0, 0): predicted thresholds on
            s.set_tuning("gemm_passes", passes)
            s.set_tuning("gemm_debug", dbg)
            s.search(Q, a.k)
            s.gemm_stats()
            reps = 20
            with Sampler() as sm:
                for _ in range(reps):
                    s.search(Q, a.k)
            st = s.gemm_stats()
            ms = st["gemm_ms"] / reps
            mhz, watts = sorted(sm.mhz), sorted(sm.watts)
            print(json.dumps({"space": space, "gemm_passes": passes, "gemm_wide": wide, "gemm_debug": dbg, "gemm_ms": round(ms, 3),
                              "TFLOPs": round(2 * a.nq * a.rows * a.dim / ms / 1e9, 1), "sm_mhz_median": mhz[len(mhz) // 2],
                              "sm_mhz_min": mhz[0], "watts_median": round(watts[len(watts) // 2]), "watts_max": round(watts[-1])}), flush=True)
    s.close()
