"""BASELINE.json configs[0..4] for bench.py's ``configs`` array (timed AFTER the headline region, same process).

Every entry: {"name", "workload", "value", "unit", ..., "roofline": {...}, "parity": {...}}.  The parity entry is an
untimed check of some of the timed queries against the streamed C oracle (oracle/exact_scan.c::orc_knn_synthetic /
orc_range_synthetic) on rank 0.  A config that fails records {"name", "error"} and the others still run.

  c1  10k x 128 cosine k=10, single query through GpuIndex.search           us / query (L2-resident: latency-bound)
  c2  1M x 768 cosine k=10, batch-1 on one GPU                               q/s, GB/s vs the HBM peak
  c3  10M x 768 l2 k=100, 4096-query batches, rows sharded over N GPUs       q/s, F_alg/t vs the TF32 peak, tier counters
  c4  10M x 384 cosine k=10, metadata filter at 1 / 10 / 50 % selectivity    q/s, B_alg/t and streamed bytes/t
  c5  100M x 128 ip: batch-1 kNN k=10 and range search (~100 hits), N GPUs   q/s, GB/s per GPU
c1, c2, c4 are single-GPU configs: rank 0 runs them while the other ranks wait.
"""
from __future__ import annotations

import json
import os
import time
import traceback

import numpy as np

SEED = 42


def _oracle():
    from oracle import cscan, exact
    cscan.use_all_cores()
    return cscan, exact


def _check_knn(rows, dists, counts, first, n, dim, scaled, Q, k, space, allow=None, to_gen=None):
    """-> parity dict for nq checked queries (rows: what the product returned, mapped to generator rows by to_gen)."""
    cscan, _ = _oracle()
    t0 = time.perf_counter()
    gen = np.stack([to_gen(r) for r in rows]) if to_gen is not None else np.asarray(rows)
    msg = cscan.check_knn_synthetic(gen, np.asarray(dists), np.asarray(counts), SEED, first, n, dim, scaled, Q, k, space, allow=allow)
    return {"queries": int(len(Q)), "ok": msg is None, "problem": msg,
            "how": f"streamed C oracle over {n} generator rows, check_topk_parity ({time.perf_counter() - t0:.1f} s)"}


def run(ctx, a, wanted, hbm_peak, bf16_peak, peak_src):
    out = []
    for name in wanted:
        fn = {"c1": c1, "c2": c2, "c3": c3, "c4": c4, "c5": c5}.get(name.strip())
        if fn is None:
            continue
        single = name in ("c1", "c2", "c4")
        entry = None
        try:
            if not single or ctx.rank == 0:
                entry = fn(ctx, a, hbm_peak, bf16_peak, peak_src)
        except Exception as e:  # noqa: BLE001
            entry = {"name": name, "error": f"{type(e).__name__}: {e}", "trace": traceback.format_exc()[-600:]}
        ctx.torch.cuda.synchronize()
        ctx.torch.cuda.empty_cache()
        ctx.barrier()
        if ctx.rank == 0 and entry is not None:
            out.append(entry)
    return out


def _scaled(n, a, floor=1000):
    return max(floor, int(n * a.configs_scale))


# ------------------------------------------------------------------------------------------------- c1
def c1(ctx, a, hbm_peak, bf16_peak, peak_src):
    from mlvectordb_b200 import GpuIndex, VectorDTO
    cscan, _ = _oracle()
    n, dim, k, reps = 10_000, 128, 10, 2000
    X = cscan.fill_synthetic(SEED, 0, n, dim, True)
    index = GpuIndex(space="cosine", device=ctx.local_rank)
    index.add_matrix(X, "c1")
    Q = np.random.default_rng(SEED + 11).standard_normal((64, dim), dtype=np.float32)
    dtos = [VectorDTO(values=Q[i], metadata={}) for i in range(64)]
    for i in range(200):
        index.search(dtos[i % 64], top_k=k, namespace="c1", metric="cosine")
    t0 = time.perf_counter()
    for i in range(reps):
        hits = index.search(dtos[i % 64], top_k=k, namespace="c1", metric="cosine")
    us = (time.perf_counter() - t0) / reps * 1e6
    shard = index._ns["c1"].shard
    t0 = time.perf_counter()
    for i in range(reps):
        shard.search(Q[i % 64][None, :], k)
    us_shard = (time.perf_counter() - t0) / reps * 1e6
    d, r, c = shard.search(Q[:4], k)
    single = [shard.search(Q[i:i + 1], k) for i in range(4)]
    same = all(np.array_equal(single[i][1][0], r[i]) and np.array_equal(single[i][0][0], d[i]) for i in range(4))
    parity = _check_knn(r, d, c, 0, n, dim, True, Q[:4], k, "cosine")
    parity["ok"] = bool(parity["ok"] and same and len(hits) == k)
    index.close()
    return {"name": "c1", "workload": f"{n}x{dim} fp32 cosine k={k}, single query (BASELINE configs[0])",
            "value": us, "unit": "us/query", "higher_is_better": False, "queries_per_sec": 1e6 / us,
            "api": "GpuIndex.search(VectorDTO, top_k, namespace, metric) -> List[SearchResult], host in / host out",
            "c_abi_us_per_query": us_shard, "c_abi": "DeviceShard.search -> mlv_index_search (host buffers)",
            "roofline": {"bound": "latency", "note": f"{n * dim * 4 / 1e6:.2f} MB matrix is L2-resident: the floor is launch + completion latency, not HBM",
                         "achieved": None, "peak": None, "frac": None},
            "parity": parity}


# ------------------------------------------------------------------------------------------------- c2
def c2(ctx, a, hbm_peak, bf16_peak, peak_src):
    torch = ctx.torch
    from mlvectordb_b200 import DeviceShard
    n, dim, k, nq = _scaled(1_000_000, a), 768, 10, 64
    s = DeviceShard(dim, "cosine", capacity=n, device=ctx.local_rank)
    s.add_synthetic(SEED, 0, n, True)
    Q = np.random.default_rng(SEED + 12).standard_normal((nq, dim), dtype=np.float32)
    Qd = torch.from_numpy(Q).to(ctx.device)
    streams = [torch.cuda.Stream(ctx.device) for _ in range(2)]
    outs = [(torch.empty((1, k), dtype=torch.float32, device=ctx.device), torch.empty((1, k), dtype=torch.int64, device=ctx.device),
             torch.empty((1,), dtype=torch.int32, device=ctx.device)) for _ in range(2)]

    def pass_(lanes):
        for j in range(nq):
            st, o = streams[j % lanes], outs[j % lanes]
            with torch.cuda.stream(st):
                s.search_device(Qd[j:j + 1].data_ptr(), 1, k, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), stream=st.cuda_stream)

    def timed(lanes, reps=3):
        pass_(lanes)
        cur = torch.cuda.current_stream(ctx.device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(cur)
        for st in streams:
            st.wait_event(e0)
        for _ in range(reps):
            pass_(lanes)
        for st in streams:
            cur.wait_stream(st)
        e1.record(cur)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (reps * nq)

    ms2, ms1 = timed(2), timed(1)
    s.search(Q[:1], k)
    t0 = time.perf_counter()
    for j in range(nq):
        s.search(Q[j:j + 1], k)
    ms_host = (time.perf_counter() - t0) / nq * 1e3
    d, r, c = s.search(Q[:4], k)
    parity = _check_knn(r, d, c, 0, n, dim, True, Q[:4], k, "cosine")
    hs = s.gemm_stats()
    s.close()
    b = n * dim * 4
    return {"name": "c2", "workload": f"{n}x{dim} fp32 cosine k={k}, batch-1 queries on one GPU (BASELINE configs[1])",
            "value": 1e3 / ms2, "unit": "queries/s", "higher_is_better": True, "queries_in_flight": 2,
            "one_query_in_flight": 1e3 / ms1, "e2e_host_buffers_qps": 1e3 / ms_host, "e2e_api": "DeviceShard.search -> mlv_index_search",
            "roofline": {"bound": "hbm", "achieved": b / (ms2 * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": b / (ms2 * 1e-3) / 1e9 / hbm_peak, "bytes_per_launch": b, "peak_source": peak_src,
                         "one_query_in_flight_GBps": b / (ms1 * 1e-3) / 1e9,
                         "shadow_scan_searches": int(hs["half_scan_queries"]), "shadow_scan_uncertified": int(hs["half_scan_uncertified"]),
                         "how": "rows*dim*4 (the fp32 matrix, SURVEY 8d) per search / (device time of the region / searches); "
                                "searches counted in shadow_scan_searches read the fp16 shadow instead (half the bytes, exact "
                                "re-rank + certificate, DESIGN 4.1b), which is why frac can exceed 1"},
            "parity": parity}


# ------------------------------------------------------------------------------------------------- c3
def c3(ctx, a, hbm_peak, bf16_peak, peak_src):
    torch = ctx.torch
    from mlvectordb_b200.sharded import ShardedIndex, shard_range
    rows, dim, k, nq, reps = _scaled(10_000_000, a, 20_000), 768, 100, 4096, 3
    idx = ShardedIndex(dim, "l2", rows, device=ctx.device)
    idx.add_synthetic(SEED, scaled=False)
    Q = np.random.default_rng(SEED + 13).uniform(-1.0, 1.0, (nq, dim)).astype(np.float32)
    Qd = torch.from_numpy(Q).to(ctx.device)
    idx.shard.set_timing(True)
    idx.search_device(Qd, k)          # warm-up: row norms, scratch, tier statistics
    idx.search_device(Qd, k)
    torch.cuda.synchronize()
    st0 = idx.shard.gemm_stats()
    res = [None]

    def batch():
        res[0] = idx.search_device(Qd, k)

    ms = ctx.timed(batch, reps) / reps
    st = idx.shard.gemm_stats()
    fast = (st["fast_queries"] - st0["fast_queries"]) / reps
    fallback = (st["fallback_queries"] - st0["fallback_queries"]) / reps
    gemm_ms = st["gemm_ms"] / reps
    idx.shard.set_timing(False)
    d, r, c = (t.cpu().numpy() for t in res[0])
    idx.search(Q, k)                   # warm-up: pinned staging is allocated by the first host-buffer call
    ctx.barrier()
    t0 = time.perf_counter()
    e2e = idx.search(Q, k)
    e2e_ms = ctx.max_over_ranks((time.perf_counter() - t0) * 1e3)
    pick = np.array([0, nq // 2 + 1, nq - 1])
    parity = None
    if ctx.rank == 0:
        parity = _check_knn(r[pick], d[pick], c[pick], 0, rows, dim, False, Q[pick], k, "l2")
        parity["ok"] = bool(parity["ok"] and all(np.array_equal(x, y) for x, y in zip(e2e, (d, r, c))))
    lo, hi = shard_range(rows, ctx.rank, ctx.world)
    idx.close()
    f_alg_gpu = 2.0 * nq * (hi - lo) * dim                       # F_alg = 2 nq N d, this GPU's rows
    tf32_peak = bf16_peak / 2
    sustained = _sustained_bf16()
    work_mult = 1.0 + 3.0 * (1.0 - fast / nq)                    # tier 1: one MMA per product; queries re-run by 3xTF32: + 3
    return {"name": "c3", "workload": f"{rows}x{dim} fp32 l2 k={k}, {nq}-query batches, rows sharded over {ctx.world} GPU(s) (BASELINE configs[2])",
            "value": nq / ms * 1e3, "unit": "queries/s", "higher_is_better": True, "ms_per_batch": ms, "n_gpus": ctx.world,
            "e2e_host_buffers_qps": nq / e2e_ms * 1e3, "e2e_api": "ShardedIndex.search(host ndarray [4096, 768], k)",
            "tiers": {"first_tier_certified_per_batch": fast, "first_tier": "fp16 shadow, CTA-pair kernel (tcgen05 cta_group::2 kind::f16)",
                      "scan_fallback_per_batch": fallback,
                      "rank0_gemm_ms_per_batch": gemm_ms, "rank0_gemm_share_of_batch": gemm_ms / ms},
            "roofline": {"bound": "tensor", "achieved": f_alg_gpu / (ms * 1e-3) / 1e12, "peak": tf32_peak, "unit": "TFLOP/s",
                         "frac": f_alg_gpu / (ms * 1e-3) / 1e12 / tf32_peak,
                         "peak_source": peak_src + " bf16_tflops / 2 (TF32 dense = half of bf16)",
                         "tensor_pipe_work_multiplier": work_mult,
                         "kernel_only_TFLOPs": (f_alg_gpu * work_mult / (gemm_ms * 1e-3) / 1e12) if gemm_ms else None,
                         "frac_of_bf16_peak": f_alg_gpu / (ms * 1e-3) / 1e12 / bf16_peak,
                         "bf16_peak_sustained": sustained,
                         "frac_of_bf16_sustained": (f_alg_gpu / (ms * 1e-3) / 1e12 / sustained) if sustained else None,
                         "how": "per GPU: F_alg = 2*nq*local_rows*dim (fp32-equivalent useful flops) / batch time; the first tier executes "
                                "1 x F_alg on the tensor pipe as kind::f16 MMAs on an fp16 shadow of the rows (its roof is the bf16/fp16 peak, "
                                "frac_of_bf16_peak; frac is against the TF32 peak north_star names), queries it cannot certify add 3 x F_alg (3xTF32); "
                                "batches run back to back at the board's power limit, so the like-for-like roof is the SUSTAINED cuBLAS "
                                "figure of MEASURED_PEAKS.json (frac_of_bf16_sustained), not the burst one"},
            "parity": parity}


# ------------------------------------------------------------------------------------------------- c4
def c4(ctx, a, hbm_peak, bf16_peak, peak_src):
    from mlvectordb_b200 import DeviceShard
    from oracle import synthetic
    n, dim, k, nq = _scaled(10_000_000, a, 20_000), 384, 10, 32
    s = DeviceShard(dim, "cosine", capacity=n, device=ctx.local_rank)
    s.add_synthetic(SEED, 0, n, True)
    buckets = synthetic.buckets(SEED + 2, 0, n)
    s.set_column(0, buckets)
    Q = np.random.default_rng(SEED + 14).standard_normal((nq, dim), dtype=np.float32)
    sel_out = []
    ok_all = True
    for pct in (1, 10, 50):
        t0 = time.perf_counter()
        f = s.where([(0, "<", pct)])
        passing = f.passing
        where_ms = (time.perf_counter() - t0) * 1e3
        s.search(Q[:1], k, f)
        s.set_timing(True)
        s.scan_time_ms()
        t0 = time.perf_counter()
        for j in range(nq):
            s.search(Q[j:j + 1], k, f)
        wall_ms = (time.perf_counter() - t0) / nq * 1e3
        scan_ms, scan_n = s.scan_time_ms()
        s.set_timing(False)
        per = scan_ms / max(scan_n, 1)
        d, r, c = s.search(Q[:3], k, f)
        par = _check_knn(r, d, c, 0, n, dim, True, Q[:3], k, "cosine", allow=buckets < pct)
        ok_all = ok_all and par["ok"]
        b_alg = passing * dim * 4 + n / 8
        sel_out.append({"selectivity_pct": pct, "passing_rows": int(passing), "queries_per_sec": 1e3 / wall_ms, "ms_per_query_host_api": wall_ms,
                        "scan_kernel_ms": per, "where_ms_first_use": where_ms,
                        "roofline": {"bound": "hbm", "achieved": b_alg / (per * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                     "frac": b_alg / (per * 1e-3) / 1e9 / hbm_peak, "bytes_per_launch": b_alg,
                                     "how": "B_alg = passing_rows*dim*4 + N/8 (SURVEY 8d) / scan kernel time (CUDA events)"},
                        "parity": par})
        f.close()
    s.search(Q[:1], k)
    s.set_timing(True)
    s.scan_time_ms()
    for j in range(8):
        s.search(Q[j:j + 1], k)
    full_ms, full_n = s.scan_time_ms()
    s.close()
    return {"name": "c4", "workload": f"{n}x{dim} fp32 cosine k={k}, metadata-filtered batch-1 kNN at 1 / 10 / 50 % selectivity on one GPU (BASELINE configs[3])",
            "value": sel_out[0]["queries_per_sec"], "unit": "queries/s", "higher_is_better": True,
            "api": "DeviceShard.where([(column, '<', v)]) -> prepared filter; DeviceShard.search(q, k, filter) (host buffers)",
            "unfiltered_scan_ms": full_ms / max(full_n, 1), "selectivities": sel_out,
            "roofline": sel_out[1]["roofline"], "parity": {"ok": bool(ok_all), "queries": 9, "how": "per selectivity, see selectivities[].parity"}}


def _sustained_bf16():
    """cuBLAS bf16 TFLOP/s held for seconds (power-limited clocks), when the driver measured it."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["bf16_tflops_sustained"])
    except (OSError, KeyError, ValueError):
        return None


# ------------------------------------------------------------------------------------------------- c5
def c5(ctx, a, hbm_peak, bf16_peak, peak_src):
    torch = ctx.torch
    from mlvectordb_b200.sharded import ShardedIndex
    cscan, exact = _oracle()
    rows, dim, k, nq = _scaled(100_000_000, a, 20_000), 128, 10, 64
    idx = ShardedIndex(dim, "ip", rows, device=ctx.device)
    idx.add_synthetic(SEED, scaled=True)
    Q = np.random.default_rng(SEED + 15).standard_normal((nq, dim), dtype=np.float32)
    Qd = torch.from_numpy(Q).to(ctx.device)
    streams = [torch.cuda.Stream(ctx.device) for _ in range(2)]
    last = [None, None]

    def knn_pass():
        for j in range(nq):
            with torch.cuda.stream(streams[j & 1]):
                last[j & 1] = idx.search_device(Qd[j:j + 1], k)

    knn_pass()
    ms = ctx.timed(knn_pass, 2, streams) / (2 * nq)
    local_bytes = (idx.hi - idx.lo) * dim * 4
    # range search through the host API: radius = each query's 100th smallest distance (~100 hits)
    nr = 16
    d100, r100, c100 = idx.search(Q[:nr], 100)
    radii = d100[:, 99].astype(np.float64)
    idx.range_search(Q[:1], float(radii[0]))
    ctx.barrier()
    t0 = time.perf_counter()
    hits = []
    for j in range(nr):
        got = idx.range_search(Q[j:j + 1], float(radii[j]))
        hits.append(got[0])
    torch.cuda.synchronize()
    range_ms = ctx.max_over_ranks((time.perf_counter() - t0) / nr * 1e3)
    d10, r10, c10 = idx.search(Q[:2], k)
    parity = None
    if ctx.rank == 0:
        parity = _check_knn(r10, d10, c10, 0, rows, dim, True, Q[:2], k, "ip")
        (ol, od), = cscan.range_synthetic(SEED, 0, rows, dim, True, Q[:1], float(radii[0]), "ip")
        hd, hr = hits[0]
        tol = 1e-5 * abs(radii[0]) + 1e-6
        odd = sorted(set(hr.tolist()) ^ set(ol.tolist()))
        odd_ok = all(abs(float(cscan.distances_synthetic(SEED, [x], dim, True, Q[0], "ip")[0]) - radii[0]) <= 2 * tol for x in odd)
        parity["range"] = {"hits": int(len(hr)), "oracle_hits": int(len(ol)), "ids_differing_at_the_radius": len(odd),
                           "ok": bool(odd_ok and abs(len(hr) - len(ol)) <= len(odd) and (np.diff(hd) >= 0).all())}
        parity["ok"] = bool(parity["ok"] and parity["range"]["ok"])
        parity["queries"] = 3
    idx.close()
    return {"name": "c5", "workload": f"{rows}x{dim} fp32 ip: batch-1 kNN k={k} (2 in flight) and range search (~100 hits/query), rows sharded over {ctx.world} GPU(s) (BASELINE configs[4])",
            "value": 1e3 / ms, "unit": "queries/s", "higher_is_better": True, "n_gpus": ctx.world,
            "range_search": {"queries_per_sec": 1e3 / range_ms, "ms_per_query_host_api": range_ms, "hits_per_query": [int(len(h[1])) for h in hits[:8]],
                             "ratio_to_knn": (1e3 / range_ms) / (1e3 / ms), "per_gpu_GBps": local_bytes / (range_ms * 1e-3) / 1e9,
                             "api": "ShardedIndex.range_search(host query, radius) -> (dists, global rows), one query at a time"},
            "roofline": {"bound": "hbm", "achieved": local_bytes / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": local_bytes / (ms * 1e-3) / 1e9 / hbm_peak, "bytes_per_launch": local_bytes, "peak_source": peak_src,
                         "how": "per GPU: local_rows*dim*4 per scan launch / (device time of the region / launches), max over ranks"},
            "parity": parity}
