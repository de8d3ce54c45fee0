"""Row-sharded configurations of BASELINE.json on N GPUs (one process per GPU, launch with torchrun):

  c3  10M x 768 l2, k=100, 4096-query batches: local tensor-core path (csrc/gemm_kernel.cuh) per
      shard, NCCL all-gather of the candidates, merge kernel
  c5  100M x 128 ip: batch-1 kNN k=10 (fused scan + peer-memory exchange, two queries in flight) and
      range search with ~100 hits per query (radius = the 100th smallest distance of each query)

Dev / profiling tool (the driver's bench is bench.py); prints one JSON line per measurement on rank 0.
Device timing: CUDA events on the launching stream, max over ranks."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlvectordb_b200.sharded import ShardedIndex  # noqa: E402
from oracle import synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c3", choices=["c3", "c5"])
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        return max_over_ranks(e0.elapsed_time(e1)), out

    def emit(d):
        if rank == 0:
            print(json.dumps(d), flush=True)

    if a.config == "c3":
        rows, dim, k, nq = a.rows or 10_000_000, 768, 100, 4096
        idx = ShardedIndex(dim, "l2", rows, device=device)
        idx.add_synthetic(42, scaled=False)
        Qd = torch.from_numpy(synthetic.queries(43, nq, dim)).to(device)
        idx.shard.set_timing(True)
        idx.search_device(Qd, k)
        torch.cuda.synchronize()
        st0 = idx.shard.gemm_stats()
        ms, _ = timed(lambda: [idx.search_device(Qd, k) for _ in range(a.reps)])
        st = idx.shard.gemm_stats()
        for key in ("fast_queries", "fallback_queries"):
            st[key] -= st0[key]
        per = ms / a.reps
        local_rows = idx.hi - idx.lo
        gms = st["gemm_ms"] / a.reps
        emit({"config": "c3", "workload": f"{rows}x{dim} l2 k={k}, {nq}-query batches, rows sharded over {world} GPU(s)",
              "n_gpus": world, "ms_per_batch": round(per, 3), "qps": round(nq / per * 1e3, 1),
              "fp32_equiv_TFLOPs_total": round(2.0 * nq * rows * dim / (per * 1e-3) / 1e12, 1),
              "rank0_gemm_ms_per_batch": round(gms, 3),
              "rank0_fast_tier_queries_per_batch": st["fast_queries"] / a.reps,
              # MMAs per product: 1 for queries the one-pass tier certified, 1 + 3 for those re-run by the 3xTF32 tier
              "rank0_tensor_TFLOPs_executed": round((1 + 3 * (1 - st["fast_queries"] / (a.reps * nq))) * 2.0 * nq
                                                    * ((local_rows + 127) // 128 * 128) * dim / (gms * 1e-3) / 1e12, 1),
              "rank0_fallback_queries": st["fallback_queries"]})
    else:
        rows, dim, k = a.rows or 100_000_000, 128, 10
        idx = ShardedIndex(dim, "ip", rows, device=device)
        idx.add_synthetic(42, scaled=True)
        nq = 64
        Q = synthetic.queries(43, nq, dim)
        Qd = torch.from_numpy(Q).to(device)
        streams = [torch.cuda.Stream(device) for _ in range(2)]

        def knn_pass():
            cur = torch.cuda.current_stream(device)
            for st in streams:
                st.wait_stream(cur)
            for j in range(nq):
                with torch.cuda.stream(streams[j & 1]):
                    idx.search_device(Qd[j:j + 1], k)
            for st in streams:
                cur.wait_stream(st)

        knn_pass()
        ms, _ = timed(lambda: [knn_pass() for _ in range(a.reps)])
        per = ms / (a.reps * nq)
        local_bytes = (idx.hi - idx.lo) * dim * 4
        emit({"config": "c5-knn", "workload": f"{rows}x{dim} ip k={k} batch-1 (2 in flight), rows sharded over {world} GPU(s)",
              "n_gpus": world, "ms_per_query": round(per, 4), "qps": round(1e3 / per, 1),
              "per_gpu_GBps": round(local_bytes / (per * 1e-3) / 1e9, 1)})
        # range search, ~100 hits per query
        d100, _, _ = idx.search(Q[:16], 100)
        radii = d100[:, 99]
        hits = []
        idx.range_search(Q[:1], float(radii[0]))
        t0 = time.perf_counter()
        for j in range(16):
            got = idx.range_search(Q[j:j + 1], float(radii[j]))
            hits.append(len(got[0][1]))
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0) / 16
        emit({"config": "c5-range", "workload": f"{rows}x{dim} ip range search, radius = each query's 100th smallest distance",
              "n_gpus": world, "ms_per_query_host_api": round(dt * 1e3, 4), "qps": round(1 / dt, 1), "hits_per_query": hits[:8],
              "per_gpu_GBps": round(local_bytes / dt / 1e9, 1)})
    idx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
