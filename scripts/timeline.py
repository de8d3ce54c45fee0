"""Per-CTA timeline of one scan launch (globaltimer stamps): where a launch's fixed overhead goes."""
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
a = ap.parse_args()
s = DeviceShard(a.dim, "cosine", capacity=a.rows)
s.add_synthetic(42, 0, a.rows, True)
Q = synthetic.queries(1, 4, a.dim)
s.set_tuning("timeline", 1)
s.set_timing(True)
for i in range(4):
    s.search(Q[i:i + 1], a.k)
    t = s.debug_timeline().astype(np.int64)
    ms, n = s.scan_time_ms()
    t0 = t[:, 0].min()
    rel = (t - t0) / 1e3  # us
    print(json.dumps({
        "rows": a.rows, "event_ms": round(ms / n, 4), "ctas": int(t.shape[0]),
        "start_us[min,med,max]": [round(float(x), 1) for x in (rel[:, 0].min(), np.median(rel[:, 0]), rel[:, 0].max())],
        "first_tile_after_start_us[min,med,max]": [round(float(x), 1) for x in ((rel[:, 1] - rel[:, 0]).min(), np.median(rel[:, 1] - rel[:, 0]), (rel[:, 1] - rel[:, 0]).max())],
        "last_tile_done_us[min,med,max]": [round(float(x), 1) for x in (rel[:, 2].min(), np.median(rel[:, 2]), rel[:, 2].max())],
        "exit_us[min,med,max]": [round(float(x), 1) for x in (rel[:, 3].min(), np.median(rel[:, 3]), rel[:, 3].max())],
    }), flush=True)
s.close()
