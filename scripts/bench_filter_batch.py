"""Filtered query BATCHES on the tensor-core path (10M x 384 cosine, k = 10): the passing rows gathered into a dense
matrix (default for selective filters) vs the masked epilogue over every row (`gather` tuning 0)."""
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
ap.add_argument("--nq", type=int, default=1024)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
s = DeviceShard(a.dim, "cosine", capacity=a.rows)
s.add_synthetic(42, 0, a.rows, True)
s.set_column(0, synthetic.buckets(44, 0, a.rows))
Q = synthetic.queries(45, a.nq, a.dim)
for cut in (1, 10, 50):
    f = s.where([(0, "<", cut)])
    ref = None
    for mode, gather in (("gathered rows", -1), ("masked epilogue over all rows", 0)):
        s.set_tuning("gather", gather)
        out = s.search(Q, a.k, f)
        st0 = s.gemm_stats()
        t0 = time.perf_counter()
        for _ in range(a.reps):
            out = s.search(Q, a.k, f)
        wall = (time.perf_counter() - t0) / a.reps
        st = s.gemm_stats()
        same = None if ref is None else bool(all(np.array_equal(x, y) for x, y in zip(out, ref)))
        ref = out
        print(json.dumps({"selectivity": cut / 100, "passing": f.passing, "nq": a.nq, "mode": mode, "ms_per_batch": round(wall * 1e3, 3),
                          "qps": round(a.nq / wall, 1), "gathered_batches": st["gathered_searches"] - st0["gathered_searches"],
                          "fallback_queries": st["fallback_queries"] - st0["fallback_queries"], "identical_to_previous_mode": same}),
              flush=True)
    s.set_tuning("gather", -1)
    f.close()
s.close()
