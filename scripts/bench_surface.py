"""The layers either side of the index, measured on one B200:

  ingest    per-row ``upsert_many`` (the reference path: one object + uuid4 per row) vs bulk ``upsert_matrix``
  search    the same batch-1 cosine k=10 queries through every layer: C ABI (``DeviceShard.search``) ->
            ``GpuIndex.search`` -> ``GpuQueryProcessor.find_similar`` (enrich on / off) -> HTTP ``POST /search``
            (in-process ASGI client; include_values on / off)
  filter    metadata constraint decided on the device columns vs on the host
  snapshot  ``GpuIndex.save`` / ``GpuIndex.load`` of the namespace

One JSON line per measurement."""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time
from collections import defaultdict

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlvectordb_b200 import GpuIndex, GpuQueryProcessor, VectorDTO  # noqa: E402
from oracle import synthetic  # noqa: E402


class DictStorage:
    """What the reference's StorageEngineInMemory does for the calls on this path (storage_engine_in_memory.py:10-86)."""

    def __init__(self):
        self._data = defaultdict(dict)

    def write_vectors(self, vectors, namespace):
        d = self._data[namespace]
        for v in vectors:
            d[v.id] = v

    def write(self, vector, namespace):
        self._data[namespace][vector.id] = vector

    def read_vectors(self, ids, namespace):
        d = self._data[namespace]
        return [d.get(i) for i in ids]

    def delete(self, vid, namespace):
        return self._data[namespace].pop(vid, None) is not None

    @property
    def namespace_map(self):
        return {ns: list(d.values()) for ns, d in self._data.items()}

    @property
    def list_namespaces(self):
        return list(self._data)

    def get_storage_info(self):
        return {"total_vectors": sum(len(d) for d in self._data.values())}


def emit(**kw):
    print(json.dumps(kw), flush=True)


def timed(fn, reps):
    fn()
    t0 = time.perf_counter()
    for i in range(reps):
        fn(i)
    return (time.perf_counter() - t0) / reps


ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=1_000_000)
ap.add_argument("--dim", type=int, default=768)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--reps", type=int, default=200)
ap.add_argument("--per-row", type=int, default=20_000)
a = ap.parse_args()

storage = DictStorage()
index = GpuIndex(space="cosine", capacity=a.rows + a.per_row)
qp = GpuQueryProcessor(storage, index)
X = synthetic.rows(42, 0, a.rows, a.dim, scaled=True)
buckets = synthetic.buckets(42, 0, a.rows)

# ---- ingest -------------------------------------------------------------------------------------------
dtos = [VectorDTO(values=X[i], metadata={"bucket": int(buckets[i])}) for i in range(a.per_row)]
t0 = time.perf_counter()
qp.upsert_many(dtos, "rowwise")
t = time.perf_counter() - t0
emit(what="ingest", path="upsert_many (one object + uuid4 per row, metadata -> device column)", rows=a.per_row,
     rows_per_s=round(a.per_row / t), GBps=round(a.per_row * a.dim * 4 / t / 1e9, 3))
mds = [{"bucket": int(b)} for b in buckets]
t0 = time.perf_counter()
ids = qp.upsert_matrix(X, "bulk", metadata=mds)
t = time.perf_counter() - t0
emit(what="ingest", path="upsert_matrix (one H2D append, ids minted in bulk, storage objects + metadata columns)", rows=a.rows,
     rows_per_s=round(a.rows / t), GBps=round(a.rows * a.dim * 4 / t / 1e9, 3))
t0 = time.perf_counter()
index.add_matrix(X, "index_only", columns={"bucket": buckets})
t = time.perf_counter() - t0
emit(what="ingest", path="GpuIndex.add_matrix (index only: H2D append + normalise + column)", rows=a.rows,
     rows_per_s=round(a.rows / t), GBps=round(a.rows * a.dim * 4 / t / 1e9, 3))

from mlvectordb_b200 import DeviceShard  # noqa: E402
for staged in (0, 1):
    sh = DeviceShard(a.dim, "cosine", capacity=a.rows)
    sh.set_tuning("staged_upload", staged)
    sh.add(X[:1000])
    t0 = time.perf_counter()
    sh.add(X[1000:])
    t = time.perf_counter() - t0
    emit(what="ingest", path="C ABI mlv_index_add, pageable host rows, " + ("two pinned chunks filled by worker threads" if staged
         else "plain cudaMemcpy2D"), rows=a.rows - 1000, rows_per_s=round((a.rows - 1000) / t), GBps=round((a.rows - 1000) * a.dim * 4 / t / 1e9, 2))
    sh.close()

# ---- search through the layers ---------------------------------------------------------------------------
Q = synthetic.queries(43, a.reps + 1, a.dim)
shard = index._ns["bulk"].shard
layers = {}
layers["C ABI: DeviceShard.search"] = lambda i=0: shard.search(Q[i:i + 1], a.k)
layers["GpuIndex.search"] = lambda i=0: index.search(VectorDTO(values=Q[i]), a.k, "bulk", "cosine")
layers["GpuQueryProcessor.find_similar(enrich=False)"] = lambda i=0: qp.find_similar(VectorDTO(values=Q[i]), a.k, "bulk", "cosine", enrich=False)
layers["GpuQueryProcessor.find_similar (reference response: values + metadata)"] = lambda i=0: qp.find_similar(VectorDTO(values=Q[i]), a.k, "bulk", "cosine")
try:
    from fastapi.testclient import TestClient
    from mlvectordb_b200.rest_api import GpuRestAPI
    client = TestClient(GpuRestAPI(qp, log_level="ERROR").get_app())
    bodies = [{"query": Q[i].tolist(), "top_k": a.k, "metric": "cosine"} for i in range(a.reps + 1)]
    layers["HTTP POST /search include_values=false (in-process ASGI)"] = lambda i=0: client.post(
        "/search", params={"namespace": "bulk"}, json=dict(bodies[i], include_values=False)).json()
    layers["HTTP POST /search (reference response)"] = lambda i=0: client.post("/search", params={"namespace": "bulk"}, json=bodies[i]).json()
except Exception as e:  # noqa: BLE001
    emit(what="search", skipped=f"REST layer: {e}")
base = None
for name, fn in layers.items():
    t = timed(fn, a.reps)
    base = base or t
    emit(what="search", layer=name, rows=a.rows, dim=a.dim, k=a.k, ms_per_query=round(t * 1e3, 4), qps=round(1 / t, 1),
         overhead_vs_c_abi_ms=round((t - base) * 1e3, 4))

# ---- filter: device columns vs host ------------------------------------------------------------------------
cons = {"bucket": ("<", 10)}
index.where("bulk", {"bucket": ("<", 11)})          # first use: allocations, kernel load
t0 = time.perf_counter()
f = index.where("bulk", cons)
t_dev = time.perf_counter() - t0
t0 = time.perf_counter()
by_id = {v.id: v.metadata for v in storage.namespace_map["bulk"]}
pf = index.prepare_filter("bulk", lambda uid: by_id[uid]["bucket"] < 10)
t_host = time.perf_counter() - t0
emit(what="filter", constraint="bucket < 10", rows=a.rows, passing=f.passing, device_where_ms=round(t_dev * 1e3, 3),
     host_predicate_ms=round(t_host * 1e3, 1), same_passing=f.passing == pf.passing)
t = timed(lambda i=0: qp.find_similar(VectorDTO(values=Q[i]), a.k, "bulk", "cosine", filter=cons, enrich=False), a.reps)
emit(what="filter", layer="find_similar(filter={'bucket': ('<', 10)}, enrich=False), filter cached", ms_per_query=round(t * 1e3, 4),
     qps=round(1 / t, 1))
pf.close()

# ---- snapshot ---------------------------------------------------------------------------------------------
tmp = tempfile.mkdtemp(prefix="mlv_snapshot_")
try:
    only = GpuIndex(space="cosine", capacity=a.rows)
    only.add_matrix(X, "bulk", columns={"bucket": buckets})
    before = only.search_batch(Q[:4], a.k, "bulk")
    t0 = time.perf_counter()
    only.save(tmp)
    t_save = time.perf_counter() - t0
    only.close()
    t0 = time.perf_counter()
    back = GpuIndex.load(tmp)
    t_load = time.perf_counter() - t0
    after = back.search_batch(Q[:4], a.k, "bulk")
    same = all(np.array_equal(x, y) for x, y in zip(before, after))
    nbytes = a.rows * a.dim * 4
    emit(what="snapshot", rows=a.rows, bytes=nbytes, save_s=round(t_save, 3), save_GBps=round(nbytes / t_save / 1e9, 2),
         load_s=round(t_load, 3), load_GBps=round(nbytes / t_load / 1e9, 2), results_identical_after_load=same, dir=tmp)
    back.close()
finally:
    shutil.rmtree(tmp, ignore_errors=True)
index.close()
