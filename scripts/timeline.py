"""Per-CTA timeline of one scan launch (globaltimer stamps): where a launch's fixed overhead goes.

    python scripts/timeline.py --rows 1250000 --dim 768                  # the 8-GPU shard of the headline
    python scripts/timeline.py --rows 10000000 --dim 384 --filter-pct 1  # gathered scan of a 1 % filter (BASELINE configs[3])
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlvectordb_b200 import DeviceShard  # noqa: E402
from oracle import synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=1_250_000)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--filter-pct", type=int, default=0)
ap.add_argument("--space", default="cosine")
a = ap.parse_args()
s = DeviceShard(a.dim, a.space, capacity=a.rows)
s.add_synthetic(42, 0, a.rows, True)
f = None
if a.filter_pct:
    s.set_column(0, synthetic.buckets(44, 0, a.rows))
    f = s.where([(0, "<", a.filter_pct)])
Q = synthetic.queries(1, 4, a.dim)
s.set_tuning("timeline", 1)
s.set_timing(True)
NAMES = {0: "start", 1: "first_tile", 2: "last_tile_done", 8: "own_sorted", 9: "all_sorted", 3: "folded", 4: "ticket"}
for i in range(4):
    s.search(Q[i:i + 1], a.k, f)
    t = s.debug_timeline().astype(np.int64)
    ms, n = s.scan_time_ms()
    t0 = t[:, 0].min()
    rel = np.where(t > 0, (t - t0) / 1e3, np.nan)  # us
    last = int(np.argmax(t[:, 5]))
    out = {"rows": a.rows, "dim": a.dim, "filter_pct": a.filter_pct, "passing": f.passing if f else a.rows,
           "event_ms": round(ms / max(n, 1), 4), "ctas": int(t.shape[0])}
    for slot, name in NAMES.items():
        out[name + "_us[min,med,max]"] = [round(float(x), 1) for x in (np.nanmin(rel[:, slot]), np.nanmedian(rel[:, slot]), np.nanmax(rel[:, slot]))]
    out["last_cta_us[ticket,fence,threshold,survivors,select_done,outputs,flag]"] = [round(float(rel[last, j]), 1) for j in (4, 10, 11, 12, 5, 6, 7)]
    print(json.dumps(out), flush=True)
s.close()
