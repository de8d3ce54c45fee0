"""Where a batch-1 search on a small (L2-resident) namespace spends its time (BASELINE configs[0]: 10k x 128 cosine k=10).

Layers, each timed over many calls with time.perf_counter:
  raw      ctypes call of mlv_index_search with preallocated buffers (the C ABI itself)
  shard    DeviceShard.search (numpy argument handling + the call)
  index    GpuIndex.search(VectorDTO) -> List[SearchResult] (uuid decoding on top)
with the one-launch latency path on and off, plus the scan kernel's own duration (CUDA events) and the per-CTA
%globaltimer timeline of one launch.  Dev / profiling tool; prints JSON lines."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlvectordb_b200 import GpuIndex, VectorDTO  # noqa: E402
from oracle import cscan  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000)
ap.add_argument("--dim", type=int, default=128)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--reps", type=int, default=3000)
a = ap.parse_args()

X = cscan.fill_synthetic(42, 0, a.rows, a.dim, True)
index = GpuIndex(space="cosine")
index.add_matrix(X, "ns")
shard = index._ns["ns"].shard
Q = np.random.default_rng(1).standard_normal((64, a.dim), dtype=np.float32)
dtos = [VectorDTO(values=Q[i], metadata={}) for i in range(64)]
lib, h = shard._lib, shard._h
d = np.empty((1, a.k), np.float32)
r = np.empty((1, a.k), np.int64)
c = np.empty(1, np.int32)


def per_call(fn, reps=a.reps):
    for i in range(200):
        fn(i)
    t0 = time.perf_counter()
    for i in range(reps):
        fn(i)
    return (time.perf_counter() - t0) / reps * 1e6


def raw(i):
    lib.mlv_index_search(h, Q[i % 64].ctypes.data, 1, a.k, None, d.ctypes.data, r.ctypes.data, c.ctypes.data)


qptrs = [Q[i].ctypes.data for i in range(64)]
dp, rp, cp = d.ctypes.data, r.ctypes.data, c.ctypes.data


def raw_cached(i):
    lib.mlv_index_search(h, qptrs[i % 64], 1, a.k, None, dp, rp, cp)


for fast in (1, 0):
    shard.set_tuning("fast_host", fast)
    out = {"rows": a.rows, "dim": a.dim, "k": a.k, "fast_host": fast,
           "raw_ctypes_cached_ptrs_us": round(per_call(raw_cached), 2),
           "raw_ctypes_us": round(per_call(raw), 2),
           "DeviceShard.search_us": round(per_call(lambda i: shard.search(Q[i % 64][None, :], a.k)), 2),
           "GpuIndex.search_us": round(per_call(lambda i: index.search(dtos[i % 64], top_k=a.k, namespace="ns", metric="cosine")), 2)}
    print(json.dumps(out), flush=True)
shard.set_tuning("fast_host", 1)
# the kernel alone (events; the timing mode uses the staged path)
shard.set_timing(True)
shard.scan_time_ms()
for i in range(200):
    shard.search(Q[i % 64][None, :], a.k)
ms, n = shard.scan_time_ms()
shard.set_timing(False)
print(json.dumps({"scan_kernel_us_events": round(ms / n * 1e3, 2), "launches": n}), flush=True)
shard.set_tuning("timeline", 1)
for fast in (0, 1, 1):
    shard.set_tuning("fast_host", fast)
    shard.search(Q[5][None, :], a.k)
    t = shard.debug_timeline().astype(np.int64)
    t0 = t[:, 0].min()
    rel = np.where(t > 0, (t - t0) / 1e3, np.nan)
    last = int(np.argmax(t[:, 5]))          # only the last CTA (ticket grid-1) stamps the final select
    names = ["start", "first_tile", "last_tile_done", "folded", "ticket"]
    out = {"fast_host": fast, "ctas": int(t.shape[0])}
    for i, nm in enumerate(names):
        out[nm + "_us[min,med,max]"] = [round(float(x), 2) for x in (np.nanmin(rel[:, i]), np.nanmedian(rel[:, i]), np.nanmax(rel[:, i]))]
    out["last_cta_us[ticket,select_done,outputs,flag]"] = [round(float(x), 2) for x in rel[last, 4:8]]
    out["last_cta_us[last_tile,own_sorted,all_sorted,pool_thr,fold_survivors,folded,ticket,fence,threshold,survivors,select_done]"] = [
        round(float(rel[last, i]), 2) for i in (2, 8, 9, 13, 14, 3, 4, 10, 11, 12, 5)]
    out["fold_survivor_count[min,med,max]"] = [int(t[:, 15].min()), int(np.median(t[:, 15])), int(t[:, 15].max())]
    print(json.dumps(out), flush=True)
shard.set_tuning("timeline", 0)
import subprocess  # noqa: E402
print(json.dumps({"clocks": subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,clocks.mem,power.draw", "--format=csv,noheader"],
                                           capture_output=True, text=True).stdout.strip()}), flush=True)
# reference points: an empty launch + sync, and a flag-poll round trip, through torch
import torch  # noqa: E402

x = torch.zeros(1, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(2000):
    x.add_(1)
    torch.cuda.synchronize()
print(json.dumps({"torch_tiny_kernel_plus_synchronize_us": round((time.perf_counter() - t0) / 2000 * 1e6, 2)}), flush=True)
index.close()
