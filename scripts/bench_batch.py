"""Batched exact kNN (BASELINE.json configs[2] shape) on one GPU: tensor-core path vs scan path.

Dev / profiling tool, not the driver's bench: prints one JSON line per configuration with the
whole-batch wall time (host queries in, host results out), the summed device time of the GEMM
launches (CUDA events inside the library), the tensor-pipe work rate

    tensor TF/s = 3 * 2 * nq_pad * rows_pad * ld_pad / gemm_time     (3xTF32: three MMAs per product)

and the fp32-equivalent F_alg = 2 * nq * rows * dim per second.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlvectordb_b200 import DeviceShard  # noqa: E402
from oracle import synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--space", default="l2")
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--nq", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--paths", default="gemm,scan")
    ap.add_argument("--scan-nq", type=int, default=256, help="queries used to time the scan path (it is slow)")
    ap.add_argument("--passes", type=int, default=0, help="gemm_passes tuning: 0 = one-pass tier then 3xTF32 (default), 1, 3")
    a = ap.parse_args()
    s = DeviceShard(a.dim, a.space, capacity=a.rows)
    t0 = time.time()
    s.add_synthetic(42, 0, a.rows, a.space != "l2")
    print(f"# filled {a.rows}x{a.dim} in {time.time() - t0:.2f}s", flush=True)
    Q = synthetic.queries(43, a.nq, a.dim)
    s.set_timing(True)
    ref = None
    for path in a.paths.split(","):
        gemm = path == "gemm"
        s.set_tuning("gemm", 1 if gemm else 0)
        s.set_tuning("gemm_passes", a.passes)
        q = Q if gemm else Q[: a.scan_nq]
        nq = q.shape[0]
        out = s.search(q, a.k)  # warm-up (allocations, row norms)
        s.gemm_stats()
        s.scan_time_ms()
        before = s.gemm_stats()
        t0 = time.perf_counter()
        for _ in range(a.reps):
            out = s.search(q, a.k)
        wall = (time.perf_counter() - t0) / a.reps
        st = s.gemm_stats()
        scan_ms, scan_n = s.scan_time_ms()
        line = {"path": path, "rows": a.rows, "dim": a.dim, "space": a.space, "k": a.k, "nq": nq,
                "wall_ms_per_batch": round(wall * 1e3, 3), "qps": round(nq / wall, 1),
                "fp32_equiv_TFLOPs": round(2.0 * nq * a.rows * a.dim / wall / 1e12, 2)}
        if gemm:
            ld = (a.dim + 3) // 4 * 4
            nq_pad = (nq + 255) // 256 * 256
            rows_pad = (a.rows + 127) // 128 * 128
            k_pad = (ld + 31) // 32 * 32
            gms = st["gemm_ms"] / a.reps
            line.update({
                "gemm_ms_per_batch": round(gms, 3),
                "rounds_per_batch": (st["rounds"] - before["rounds"]) / a.reps,
                "fallback_queries_per_batch": (st["fallback_queries"] - before["fallback_queries"]) / a.reps,
                "gemm_passes": a.passes,
                "fast_tier_queries_per_batch": (st["fast_queries"] - before["fast_queries"]) / a.reps,
                # tensor-pipe work of the launches: MMAs per product x 2 nq rows d (padded); with both tiers the
                # second tier's share depends on how many queries it re-ran, so only the pure modes are quoted
                "tensor_TFLOPs_3xTF32": round(3 * 2.0 * nq_pad * rows_pad * k_pad / (gms * 1e-3) / 1e12, 1) if a.passes == 3 else None,
                "tensor_TFLOPs_1xTF32": round(2.0 * nq_pad * rows_pad * k_pad / (gms * 1e-3) / 1e12, 1)
                if (st["fast_queries"] - before["fast_queries"]) == nq * a.reps else None,
                "gemm_share_of_wall": round(gms / (wall * 1e3), 3),
                "fallback_scan_ms": round(scan_ms / a.reps, 3),
            })
            ref = out
        else:
            line.update({"scan_ms_per_batch": round(scan_ms / a.reps, 3), "scan_passes": scan_n / a.reps})
            if ref is not None:
                same = all(np.array_equal(x[:nq], y, equal_nan=True) for x, y in zip(ref, out))
                line["identical_to_gemm_path"] = bool(same)
        print(json.dumps(line), flush=True)
    s.close()


if __name__ == "__main__":
    main()
