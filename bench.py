#!/usr/bin/env python
"""bench.py -- exact kNN queries/sec on the BASELINE.json headline workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json metric): 10M x 768 fp32 rows, cosine, k=10, batch-1 queries; with N GPUs
the rows are sharded over the ranks (strong scaling: total work fixed) and every query ends with
an NCCL all-gather of the ranks' candidates + the final select kernel.  A "step" is
``--queries-per-step`` consecutive batch-1 queries (each a full pass over the stored rows).

One JSON line on rank 0:
  value     queries/s with queries and results resident in HBM (device pointers through the C ABI,
            CUDA-event timed, max over ranks).  Every query is its own search (own launch, own
            result); two are in flight on two streams so one query's scan fills the SMs the
            previous one's tail has left (roofline.one_query_in_flight has the strictly serial rate)
  e2e       the same through the public host API: host query in, host results out, copies inside,
            two requests in flight (``GpuIndex.search_async`` at N=1, ``ShardedIndex.search_async`` at
            N>1); e2e.one_request_at_a_time is the plain synchronous ``search`` call
  roofline  scan kernel: algorithmic bytes per launch / average launch duration over the timed
            region (region device time / scan launches in it) vs the measured HBM copy peak;
            kernel_alone = CUDA events around every scan launch with one query in flight
  cpu_baseline  the oracle's C port of hnswlib's brute-force arithmetic on the host cores, on a
            bounded row sample, scaled to the full row count (N=1, rank 0 only)

``--impl reference`` times that CPU implementation as the reference arm (the reference's own
search is the absent hnswlib wheel) and prints the same line with impl=reference.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "exact_knn_queries_per_sec"
UNIT = "queries/s"
SEED = 42


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--space", default="cosine")
    ap.add_argument("--queries-per-step", type=int, default=32)
    ap.add_argument("--inflight", type=int, default=2,
                    help="independent batch-1 queries in flight (one CUDA stream each); 1 = strictly one at a time")
    ap.add_argument("--cpu-sample-rows", type=int, default=1_000_000)
    ap.add_argument("--cpu-queries", type=int, default=48)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_config(a, n_gpus):
    return {
        "workload": f"{a.rows}x{a.dim} fp32 {a.space} exact kNN k={a.k}, batch-1 queries "
                    f"(BASELINE.json metric config; rows sharded over {n_gpus} GPU(s))",
        "rows": a.rows, "dim": a.dim, "k": a.k, "space": a.space, "batch": 1,
        "queries_per_step": a.queries_per_step, "queries_in_flight": a.inflight,
        "parallelism": f"row-shard x{n_gpus}" if n_gpus > 1 else "single GPU",
        "l2": f"no flush needed: every query streams {a.rows * a.dim * 4 / n_gpus / 1e9:.2f} GB per GPU (> 126 MB L2)",
    }


def make_queries(nq, dim):
    return np.random.default_rng(SEED + 1).standard_normal((nq, dim), dtype=np.float32)


# --------------------------------------------------------------------------- clocks sampler
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index = gpu_index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.error = None

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu_index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self._stop_evt.is_set():
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
                time.sleep(0.05)
        except Exception as e:  # noqa: BLE001
            self.error = repr(e)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)

    def summary(self):
        out = {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(self.samples)}
        if self.error:
            out["error"] = self.error
        return out


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except (ValueError, IndexError):
            pass
    return local_rank


# --------------------------------------------------------------------------- CPU arm (oracle)
def cpu_knn_qps(a, steps, warmup, queries_per_step):
    """Oracle C port (hnswlib brute-force arithmetic, SIMD16 summation order, all host threads)
    on a bounded prefix of the rows; queries/s scaled linearly to the full row count."""
    from oracle import cscan

    n = min(a.cpu_sample_rows, a.rows)
    scaled = True
    X = cscan.fill_synthetic(SEED, 0, n, a.dim, scaled)
    if a.space == "cosine":
        X = cscan.normalize(X)
    Q = make_queries(queries_per_step, a.dim)
    if a.space == "cosine":
        Q = cscan.normalize(Q)
    for _ in range(warmup):
        cscan.knn(X, Q[:2], a.k, a.space)
    t0 = time.perf_counter()
    for _ in range(steps):
        for j in range(queries_per_step):
            cscan.knn(X, Q[j:j + 1], a.k, a.space)
    dt = time.perf_counter() - t0
    qps_sample = steps * queries_per_step / dt
    frac = n / a.rows
    return {
        "value": qps_sample * frac, "unit": UNIT, "cores": cscan.num_threads(), "kind": "port",
        "sample": f"first {n} of {a.rows} rows, {steps * queries_per_step} batch-1 queries in {dt:.1f} s "
                  f"({qps_sample:.2f} q/s on the sample); scaled by {frac:g} (a scan is linear in rows)",
        "impl": "oracle/exact_scan.c orc_knn (hnswlib 0.8.0 arithmetic restated; hnswlib itself is not installable here)",
    }, dt


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded: ~1.5 s of CPU work per step at 1M x 768
    qps_step = max(1, min(a.queries_per_step, 8))
    base, dt = cpu_knn_qps(a, a.steps, max(a.warmup, 1), qps_step)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {**workload_config(a, a.gpus), "queries_per_step": qps_step},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_result(line)


# --------------------------------------------------------------------------- GPU arm
def run_ours(a):
    import torch
    import torch.distributed as dist

    from mlvectordb_b200 import GpuIndex, VectorDTO
    from mlvectordb_b200.sharded import ShardedIndex

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    n_gpus = world

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- data: generated on the device, rows sharded over ranks ------------------------------
    ns = "bench"
    if world == 1:
        index = GpuIndex(space=a.space, device=local_rank, capacity=a.rows)
        index.add_synthetic(ns, a.rows, a.dim, SEED, scaled=True)
        shard = index._ns[ns].shard
        sharded = None
        local_rows = a.rows
    else:
        sharded = ShardedIndex(a.dim, a.space, a.rows, device=device)
        sharded.add_synthetic(SEED, scaled=True)
        shard = sharded.shard
        index = None
        local_rows = sharded.hi - sharded.lo
    qps_step = a.queries_per_step
    Q = make_queries(qps_step, a.dim)
    Qd = torch.from_numpy(Q).to(device)
    k = a.k
    inflight = max(1, min(a.inflight, 2))
    streams = [torch.cuda.Stream(device) for _ in range(inflight)]
    outs = [(torch.empty((1, k), dtype=torch.float32, device=device), torch.empty((1, k), dtype=torch.int64, device=device),
             torch.empty((1,), dtype=torch.int32, device=device)) for _ in range(inflight)]
    last_out = [None]

    def step_device(lanes=inflight):
        # every query is its own search (own launch, own result); `lanes` of them are in flight, each
        # on its own stream, so one query's scan fills the SMs the previous one's tail has left
        for j in range(qps_step):
            st = streams[j % lanes]
            with torch.cuda.stream(st):
                if sharded is None:
                    o = outs[j % lanes]
                    shard.search_device(Qd[j:j + 1].data_ptr(), 1, k, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(),
                                        stream=st.cuda_stream)
                    last_out[0] = o
                else:
                    last_out[0] = sharded.search_device(Qd[j:j + 1], k)

    def step_e2e(depth=inflight):
        # host query in, host results out, per query; `depth` requests in flight through the asynchronous
        # public API (search_async / .result()), 1 = the plain synchronous search() call
        last, pending = None, []
        for j in range(qps_step):
            if depth <= 1:
                if sharded is None:
                    last = index.search(VectorDTO(values=Q[j], metadata={}), top_k=k, namespace=ns, metric=a.space)
                else:
                    last = sharded.search(Q[j:j + 1], k)
                continue
            if sharded is None:
                pending.append(index.search_async(VectorDTO(values=Q[j], metadata={}), top_k=k, namespace=ns, metric=a.space))
            else:
                pending.append(sharded.search_async(Q[j:j + 1], k))
            if len(pending) >= depth:
                last = pending.pop(0).result()
        for p in pending:
            last = p.result()
        return last

    def timed(fn, n):
        """n calls of fn between two events on the default stream that every lane stream is fenced by."""
        cur = torch.cuda.current_stream(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        torch.cuda.synchronize()
        e0.record(cur)
        for st in streams:
            st.wait_event(e0)
        for _ in range(n):
            fn()
        for st in streams:
            cur.wait_stream(st)
        e1.record(cur)
        torch.cuda.synchronize()
        barrier()
        return e0.elapsed_time(e1)

    # ---- device-resident timing ----------------------------------------------------------------
    for _ in range(a.warmup):
        step_device()
    torch.cuda.synchronize()
    launches0 = shard.kernel_launches()
    merges0 = sharded.merge_launches if sharded else 0
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    ms_local = timed(step_device, a.steps)
    sampler.stop()
    ms_total = max_over_ranks(ms_local)
    launches = shard.kernel_launches() - launches0 + ((sharded.merge_launches - merges0) if sharded else 0)
    total_launches = int(sum_over_ranks(launches))
    n_queries = a.steps * qps_step
    value = n_queries / (ms_total / 1e3)
    # the scan kernel alone: one query at a time, CUDA events around every scan launch
    shard.set_timing(True)
    shard.scan_time_ms()
    alone_steps = max(1, min(a.steps, 3))
    alone_ms = timed(lambda: step_device(1), alone_steps)
    scan_ms, scan_n = shard.scan_time_ms()
    shard.set_timing(False)

    # ---- end-to-end timing through the public host API -----------------------------------------
    def time_e2e(depth):
        for _ in range(a.warmup):
            step_e2e(depth)
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            out = step_e2e(depth)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        barrier()
        return n_queries / dt, out

    e2e_value, last = time_e2e(inflight)
    e2e_sync_value, last_sync = time_e2e(1)
    assert last is not None and len(last) > 0 and last_sync is not None and len(last_sync) > 0

    # ---- roofline of the dominant kernel ---------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        with open(peaks_path) as f:
            peak, peak_src = float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    bytes_per_launch = local_rows * a.dim * 4          # R*d*4; rows pre-normalised, no bitmap
    # average launch duration over the timed region: queries overlap (2 in flight), so it is the
    # region's device time divided by the scan launches in it -- fold, final select, the multi-GPU
    # exchange wait and every launch gap included
    launch_ms = ms_total / n_queries
    achieved = bytes_per_launch / (launch_ms * 1e-3) / 1e9
    scan_ms_avg = scan_ms / max(scan_n, 1)
    traffic = None
    ncu_path = os.path.join(ROOT, "profiles", "scan_ncu_summary.json")
    if os.path.exists(ncu_path):
        try:
            with open(ncu_path) as f:
                traffic = json.load(f).get("dram_bytes_per_launch_at_bench_shape", {}).get(str(local_rows * a.dim * 4))
        except Exception:  # noqa: BLE001
            traffic = None
    roofline = {
        "bound": "hbm", "kernel": "mlv::scan_kernel<ip,NQ=1,R> (TMA ring scan + fused top-k, final select"
                                  + (", peer-memory exchange" if sharded is not None else "") + ")",
        "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "peak_source": peak_src, "traffic": traffic,
        "bytes_per_launch": bytes_per_launch, "launches_timed": n_queries, "mean_launch_ms": launch_ms,
        "how": "algorithmic bytes per scan launch / (timed-region device time / scan launches in it)",
        "frac_of_nominal_8000": achieved / 8000.0,
        "one_query_in_flight": {"value": alone_steps * qps_step / (max_over_ranks(alone_ms) / 1e3), "unit": UNIT,
                                "how": "the same searches strictly one after another on one stream (device-timed)"},
        "kernel_alone": {"mean_launch_ms": scan_ms_avg, "GBps": (bytes_per_launch / (scan_ms_avg * 1e-3) / 1e9) if scan_n else None,
                         "launches": scan_n, "share_of_its_step": (scan_ms / alone_ms) if scan_n else None,
                         "how": "CUDA events around every scan launch, one query in flight"},
    }

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(a, n_gpus),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": qps_step * a.dim * 4,
                "d2h_bytes_per_step": qps_step * (k * 12 + 4),
                "api": ("GpuIndex.search_async(VectorDTO, top_k, namespace, metric).result()" if sharded is None
                        else "ShardedIndex.search_async(host ndarray, k).result()") + f", {inflight} requests in flight",
                "one_request_at_a_time": {"value": e2e_sync_value, "unit": UNIT,
                                          "api": "GpuIndex.search(...)" if sharded is None else "ShardedIndex.search(...)"}},
        "roofline": roofline, "clocks": sampler.summary(), "gpu_launches": total_launches,
    }
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        base, _ = cpu_knn_qps(a, 1, 1, a.cpu_queries)
        line["cpu_baseline"] = base
    if rank == 0:
        emit_result(line)
    if index is not None:
        index.close()
    if sharded is not None:
        sharded.close()
    if world > 1:
        dist.destroy_process_group()


_RESULT_FD = None


def emit_result(line: dict) -> None:
    """THE one JSON line of the contract, on the process's original stdout."""
    text = json.dumps(line) + "\n"
    if _RESULT_FD is None:
        sys.stdout.write(text)
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, text.encode())


def main():
    # Libraries below us write to file descriptor 1 on their own (NCCL prints "NCCL version ..." there when
    # NCCL_DEBUG=VERSION is set in the environment): keep stdout for the result line, send the rest to stderr.
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
