"""Filtered kNN (BASELINE.json configs[3] shape: 10M x 384 cosine k=10, selectivity 1 / 10 / 50 %).

Per selectivity: batch-1 search time with (a) the per-call bitmap (gather list rebuilt every call),
(b) a prepared filter (list resident), (c) stream + mask (every row read).  Algorithmic bytes
(SURVEY.md 8d): R*d*4 with R = passing rows (+ N/8 bitmap); reports B_alg / t and streamed bytes / t."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlvectordb_b200 import DeviceShard  # noqa: E402
from oracle import synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--dim", type=int, default=384)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--sel", default="0.01,0.1,0.5")
a = ap.parse_args()
s = DeviceShard(a.dim, "cosine", capacity=a.rows)
s.add_synthetic(42, 0, a.rows, True)
Q = synthetic.queries(43, a.reps, a.dim)
bucket = np.random.default_rng([44, 0]).integers(0, 100, a.rows)
s.set_timing(True)
# predicate evaluation: metadata column on the device (mlv_filter_create_where) vs the host building the bitmap
s.set_column(0, bucket.astype(np.int32))
for sel in (float(x) for x in a.sel.split(",")):
    cut = int(round(sel * 100))
    f = s.where([(0, "<", cut)])            # warm-up (allocations)
    f.close()
    t0 = time.perf_counter()
    for _ in range(5):
        f = s.where([(0, "<", cut)])
        f.close()
    dev_ms = (time.perf_counter() - t0) / 5 * 1e3
    t0 = time.perf_counter()
    pf = s.prepare_filter(bucket < cut)
    host_ms = (time.perf_counter() - t0) * 1e3
    f = s.where([(0, "<", cut)])
    same = bool(np.array_equal(f.bitmap(), pf.bitmap())) and f.passing == pf.passing
    print(json.dumps({"selectivity": sel, "mode": "build filter from predicate bucket < %d" % cut, "passing": f.passing,
                      "device_where_ms": round(dev_ms, 3), "host_mask_upload_ms": round(host_ms, 3),
                      "bitmaps_identical": same, "column_bytes": a.rows * 4}), flush=True)
    f.close()
    pf.close()
for sel in (float(x) for x in a.sel.split(",")):
    mask = bucket < sel * 100
    passing = int(mask.sum())
    words = np.packbits(mask, bitorder="little").view(np.uint32) if a.rows % 32 == 0 else mask
    pf = s.prepare_filter(mask)
    for name, filt, gather in (("per-call bitmap, gather", words, 1), ("prepared filter, gather", pf, 1),
                               ("prepared filter, auto", pf, -1), ("stream + mask", pf, 0)):
        s.set_tuning("gather", gather)
        s.search(Q[:1], a.k, filt)
        s.scan_time_ms()
        t0 = time.perf_counter()
        for i in range(a.reps):
            s.search(Q[i:i + 1], a.k, filt)
        wall = (time.perf_counter() - t0) / a.reps
        ms, n = s.scan_time_ms()
        per = ms / n
        alg = passing * a.dim * 4 + a.rows / 8
        print(json.dumps({"selectivity": sel, "passing": passing, "mode": name, "scan_ms": round(per, 4),
                          "wall_ms_per_query": round(wall * 1e3, 4), "qps": round(1 / wall, 1),
                          "B_alg_GBps": round(alg / per / 1e6, 1),
                          "full_matrix_GBps_equiv": round(a.rows * a.dim * 4 / per / 1e6, 1)}), flush=True)
    pf.close()
s.close()
