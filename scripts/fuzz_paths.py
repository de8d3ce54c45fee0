"""Randomised differential test of the search paths on a GPU (dev tool; the deterministic cases live in tests/).

For --seconds S it draws random shapes (rows, dim, metric, k, batch width, tombstones, filters) and checks that
every way the library can answer the same question returns the same BITS as the exact scan:

  * tensor-core path in a random tier / kernel configuration (gemm_passes 0-3, gemm_wide 0/3, gemm_predict 0/1) == scan path
  * shadow scan (fp16 shadow + exact re-rank + certificate, FMA or tensor-core consumers) == fp32 scan
  * one-launch latency path (single query) == staged path == row of a batch
  * gathered filter == stream + mask filter == per-call bitmap
  * range search at the k-th distance contains the kNN answer
and, on small shapes, that the scan agrees with the CPU oracle (oracle.exact.check_topk_parity).
Prints one JSON line per failure and a summary line; exit code 1 when anything differed.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlvectordb_b200 import DeviceShard  # noqa: E402
from oracle import exact, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=120)
ap.add_argument("--seed", type=int, default=1)
a = ap.parse_args()
rng = np.random.default_rng(a.seed)
t_end = time.time() + a.seconds
cases = failures = 0


def same(x, y):
    return all(np.array_equal(p, q, equal_nan=True) for p, q in zip(x, y))


def fail(what, **ctx):
    global failures
    failures += 1
    print(json.dumps({"FAIL": what, **{k: (int(v) if isinstance(v, (np.integer,)) else v) for k, v in ctx.items()}}), flush=True)


while time.time() < t_end:
    cases += 1
    space = str(rng.choice(["l2", "ip", "cosine"]))
    dim = int(rng.choice([4, 7, 32, 33, 64, 100, 128, 200, 384, 768, 1000, 1536]))
    n = int(rng.choice([1, 50, 1000, 5000, 20_000, 70_001, 200_000]))
    if n * dim > 120_000_000:
        n = 120_000_000 // dim
    k = int(rng.choice([1, 5, 10, 27, 32, 33, 100, 200]))
    nq = int(rng.choice([1, 2, 4, 5, 9, 40, 129, 300, 600]))
    scaled = bool(rng.integers(0, 2))
    magnitude = float(rng.choice([1.0, 1.0, 1e-3, 3e4]))
    ctx = dict(space=space, dim=dim, n=n, k=k, nq=nq, scaled=scaled, magnitude=magnitude, case=cases)
    s = DeviceShard(dim, space, capacity=n)
    s.add_synthetic(1000 + cases, 0, n, scaled)
    if magnitude != 1.0 and n <= 20_000:      # re-add scaled rows through the host path (fp16 shadow scales, norms)
        X = synthetic.rows(1000 + cases, 0, n, dim, scaled) * np.float32(magnitude)
        s.clear()
        s.add(X)
    Q = synthetic.queries(2000 + cases, nq, dim) * np.float32(magnitude if rng.random() < 0.5 else 1.0)
    if n > 3:
        Q[0] = s.get_rows(np.array([n // 2], dtype=np.uint64))[0]     # a stored row (already normalised for cosine: fine)
    if rng.random() < 0.5 and n > 10:
        s.mark_deleted(rng.choice(n, size=max(1, n // 7), replace=False).astype(np.uint64))
    filt = None
    if rng.random() < 0.4 and n > 10:
        filt = rng.random(n) < float(rng.choice([0.02, 0.3, 0.9]))
    try:
        s.set_tuning("gemm", 0)
        ref = s.search(Q, k, filt)
        # tensor-core path, random configuration
        if dim >= 32 and nq >= 2:
            s.set_tuning("gemm", 1)
            passes, wide = int(rng.integers(0, 4)), int(rng.choice([0, 3]))
            s.set_tuning("gemm_passes", passes)
            s.set_tuning("gemm_wide", wide)
            s.set_tuning("gemm_predict", int(rng.integers(0, 2)))
            got = s.search(Q, k, filt)
            if not same(got, ref):
                fail("tensor-core path != scan", passes=passes, wide=wide, filtered=filt is not None, **ctx)
            s.set_tuning("gemm", 0)
            s.set_tuning("gemm_passes", 0)
            s.set_tuning("gemm_wide", 3)
        # shadow scan (single query, k <= 16) == fp32 scan
        if k <= 16 and dim >= 8:
            s.set_tuning("scan_half", 1)
            s.set_tuning("scan_half_mma", int(rng.integers(0, 2)))
            one = s.search(Q[:1], k, filt)
            s.set_tuning("scan_half", 0)
            if not same(one, tuple(x[:1] for x in ref)):
                fail("shadow scan != fp32 scan", filtered=filt is not None, **ctx)
        # latency path == staged path == batch row
        if filt is None:
            fast = s.search(Q[:1], k)
            s.set_tuning("fast_host", 0)
            staged = s.search(Q[:1], k)
            s.set_tuning("fast_host", 1)
            if not same(fast, staged):
                fail("latency path != staged path", **ctx)
            if not (np.array_equal(fast[1][0], ref[1][0]) and np.array_equal(fast[0][0], ref[0][0], equal_nan=True)):
                fail("single query != row 0 of the batch", **ctx)
        else:
            pf = s.prepare_filter(filt)
            s.set_tuning("gather", 1)
            g = s.search(Q[:4], k, pf)
            s.set_tuning("gather", 0)
            m = s.search(Q[:4], k, pf)
            s.set_tuning("gather", -1)
            if not (same(g, m) and same(g, tuple(x[:4] for x in ref))):
                fail("gather / mask / per-call bitmap disagree", **ctx)
            one = s.search(Q[:1], k, pf)           # latency path with a bound prepared filter
            if not same(one, tuple(x[:1] for x in ref)):
                fail("latency path with a prepared filter != batch", **ctx)
            pf.close()
        # range search at the k-th distance
        c0 = int(ref[2][0])
        if c0 > 0 and filt is None:
            (hd, hr), = s.range_search(Q[:1], float(ref[0][0, c0 - 1]))
            if not (len(hr) >= c0 and hr[:c0 - 1].tolist() == ref[1][0, :c0 - 1].tolist()):
                fail("range search does not contain the kNN answer", hits=len(hr), **ctx)
        # oracle on small shapes
        if n <= 5000 and magnitude == 1.0:
            X = s.get_rows(np.arange(n, dtype=np.uint64))
            info = s.info()
            live = np.ones(n, bool)
            if info.live != info.rows:
                live = np.unpackbits(s.export_live().view(np.uint8), bitorder="little")[:n].astype(bool)
            allow = live if filt is None else (live & filt)
            Qn = exact.normalize_rows(Q[:3]) if space == "cosine" else Q[:3]
            L, D = exact.knn(X, Qn, k, "ip" if space == "cosine" else space, allow=allow)   # rows come back already normalised
            # ip on unnormalised data: 1 - dot cancels, so the fp32 summation-order noise of the DOT (not of the score)
            # is what two correct implementations differ by: absolute tolerance 4 eps |x|max |q| on top of the rule's
            xmax = float(np.sqrt((X.astype(np.float64) ** 2).sum(1)).max()) if n else 0.0
            for i in range(min(3, nq)):
                c = int(ref[2][i])
                atol = exact.ATOL + (4 * 1.2e-7 * xmax * float(np.linalg.norm(Qn[i].astype(np.float64))) if space == "ip" else 0.0)
                msg = exact.check_topk_parity(ref[1][i, :c], ref[0][i, :c], L[i], D[i], atol=atol) if c == len(L[i]) else f"count {c} vs {len(L[i])}"
                if msg:
                    fail("scan != oracle", msg=msg, query=i, **ctx)
    except Exception as e:  # noqa: BLE001
        fail("exception", error=f"{type(e).__name__}: {e}"[:300], **ctx)
    s.close()
print(json.dumps({"cases": cases, "failures": failures, "seconds": a.seconds, "seed": a.seed}), flush=True)
sys.exit(1 if failures else 0)
