"""Batch-width x matrix-size surface of the two search paths on one GPU (dev tool): queries/s of the tensor-core path and
of the scan path (8 queries per pass) for the same host batch, and which one the library picks on its own."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlvectordb_b200 import DeviceShard  # noqa: E402
from oracle import synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--space", default="cosine")
ap.add_argument("--rows", default="100000,1000000,10000000")
ap.add_argument("--nq", default="8,16,64,256,1024,4096")
a = ap.parse_args()
for rows in (int(x) for x in a.rows.split(",")):
    s = DeviceShard(a.dim, a.space, capacity=rows)
    s.add_synthetic(42, 0, rows, True)
    for nq in (int(x) for x in a.nq.split(",")):
        Q = synthetic.queries(43, nq, a.dim)
        line = {"rows": rows, "dim": a.dim, "k": a.k, "nq": nq}
        for name, gemm in (("gemm", 1), ("scan", 0), ("auto", -1)):
            if name == "scan" and nq * rows > 64 * 10_000_000:
                continue
            s.set_tuning("gemm", gemm)
            before = s.gemm_stats()["searches"]
            s.search(Q, a.k)
            took_gemm = s.gemm_stats()["searches"] > before
            reps = max(2, min(20, int(2e9 / (rows * max(nq, 8)))))
            t0 = time.perf_counter()
            for _ in range(reps):
                s.search(Q, a.k)
            dt = (time.perf_counter() - t0) / reps
            line[name + "_qps"] = round(nq / dt, 1)
            line[name + "_ms"] = round(dt * 1e3, 3)
            if name == "auto":
                line["auto_path"] = "gemm" if took_gemm else "scan"
        print(json.dumps(line), flush=True)
    s.close()
