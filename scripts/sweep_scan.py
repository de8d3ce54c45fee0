"""Sweep the scan kernel's tuning knobs on one GPU and print achieved GB/s (dev tool, not the bench)."""
import argparse
import itertools
import json
import sys
import time
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlvectordb_b200 import DeviceShard  # noqa: E402
from oracle import synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--space", default="cosine")
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--nq", type=int, default=1)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--cw", default="0")
    ap.add_argument("--stage_kb", default="0")
    ap.add_argument("--evict", default="-1")
    ap.add_argument("--r", default="0")
    ap.add_argument("--max_stages", default="8")
    ap.add_argument("--ctas", default="0")
    a = ap.parse_args()
    s = DeviceShard(a.dim, a.space, capacity=a.rows)
    t0 = time.time()
    s.add_synthetic(42, 0, a.rows, True)
    print(f"# filled {a.rows}x{a.dim} in {time.time()-t0:.2f}s", flush=True)
    Q = synthetic.queries(42, max(a.nq, 1), a.dim)
    s.set_timing(True)
    bytes_alg = a.rows * a.dim * 4
    ints = lambda v: [int(x) for x in str(v).split(",")]
    for cw, skb, ev, r, ms_, ctas in itertools.product(ints(a.cw), ints(a.stage_kb), ints(a.evict), ints(a.r),
                                                      ints(a.max_stages), ints(a.ctas)):
        for key, val in (("cw", cw), ("stage_kb", skb), ("evict_first", ev), ("r", r), ("max_stages", ms_), ("ctas", ctas)):
            s.set_tuning(key, val)
        try:
            for _ in range(3):
                s.search(Q, a.k)
            s.scan_time_ms()
            t0 = time.perf_counter()
            for _ in range(a.reps):
                s.search(Q, a.k)
            wall = (time.perf_counter() - t0) / a.reps
            ms, n = s.scan_time_ms()
            per = ms / n
            passes = n / a.reps
            print(json.dumps({"cw": cw, "stage_kb": skb, "evict": ev, "r": r, "max_stages": ms_, "ctas": ctas,
                              "scan_ms": round(per, 4), "GBps": round(bytes_alg / per / 1e6, 1),
                              "passes_per_search": passes, "wall_ms_per_search": round(wall * 1e3, 4),
                              "qps": round(a.nq / wall, 1)}), flush=True)
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"cw": cw, "stage_kb": skb, "error": str(e)}), flush=True)
    s.close()


if __name__ == "__main__":
    main()
