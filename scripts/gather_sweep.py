"""Sweep of the gathered (filtered) scan's launch shape at low selectivity (dev tool; BASELINE configs[3] shape).

One factor at a time around the defaults, then the cross of the factors that moved: prints one JSON line per shape with
the scan kernel's CUDA-event time and B_alg / t (B_alg = passing rows * dim * 4 + N / 8)."""
import argparse
import itertools
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlvectordb_b200 import DeviceShard  # noqa: E402
from oracle import synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--dim", type=int, default=384)
ap.add_argument("--reps", type=int, default=40)
ap.add_argument("--sel", default="1,10")
a = ap.parse_args()
s = DeviceShard(a.dim, "cosine", capacity=a.rows)
s.add_synthetic(42, 0, a.rows, True)
Q = synthetic.queries(43, a.reps, a.dim)
s.set_column(0, synthetic.buckets(44, 0, a.rows))
s.set_timing(True)
DEFAULTS = {"pw": 0, "r": 0, "cw": 0, "stage_kb": 0, "max_stages": 8, "tile_batch": 4, "dynamic": 1, "ctas": 0}


def run(f, passing, **kw):
    for key, v in {**DEFAULTS, **kw}.items():
        s.set_tuning(key, v)
    try:
        s.search(Q[:1], 10, f)
        s.scan_time_ms()
        for i in range(a.reps):
            s.search(Q[i:i + 1], 10, f)
        ms, n = s.scan_time_ms()
    except Exception as e:  # noqa: BLE001
        return {"error": str(e)[:100], **kw}
    per = ms / n
    return {**kw, "scan_us": round(per * 1e3, 2), "B_alg_GBps": round((passing * a.dim * 4 + a.rows / 8) / per / 1e6, 1)}


for pct in (int(x) for x in a.sel.split(",")):
    f = s.where([(0, "<", pct)])
    base = run(f, f.passing)
    print(json.dumps({"selectivity_pct": pct, "passing": f.passing, "shape": "defaults", **base}), flush=True)
    for key, values in (("pw", (1, 2, 4)), ("r", (1, 2, 4)), ("cw", (4, 8, 12, 16)), ("stage_kb", (12, 24, 48, 96)),
                        ("max_stages", (2, 4, 8, 16)), ("tile_batch", (1, 2, 4)), ("dynamic", (0, 1)), ("ctas", (148, 296))):
        for v in values:
            print(json.dumps({"selectivity_pct": pct, **run(f, f.passing, **{key: v})}), flush=True)
    for pw, cw, kb in itertools.product((2, 4), (8, 16), (12, 24, 48)):
        print(json.dumps({"selectivity_pct": pct, **run(f, f.passing, pw=pw, cw=cw, stage_kb=kb, max_stages=16)}), flush=True)
    f.close()
s.close()
