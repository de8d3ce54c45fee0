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
for i in range(3):
    shard.search(Q[i][None, :], a.k)
    t = shard.debug_timeline().astype(np.int64)
    rel = (t - t[:, 0].min()) / 1e3
    print(json.dumps({"ctas": int(t.shape[0]),
                      "start_us[min,med,max]": [round(float(x), 1) for x in (rel[:, 0].min(), np.median(rel[:, 0]), rel[:, 0].max())],
                      "first_tile_us[min,med,max]": [round(float(x), 1) for x in (rel[:, 1].min(), np.median(rel[:, 1]), rel[:, 1].max())],
                      "last_tile_done_us[min,med,max]": [round(float(x), 1) for x in (rel[:, 2].min(), np.median(rel[:, 2]), rel[:, 2].max())],
                      "exit_us[min,med,max]": [round(float(x), 1) for x in (rel[:, 3].min(), np.median(rel[:, 3]), rel[:, 3].max())]}), flush=True)
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
