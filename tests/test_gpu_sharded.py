"""Multi-GPU parity: row shards on 2 (and 4/8 when present) GPUs, one process per GPU, vs the
unsharded search on one GPU -- bit-identical rows, scores and counts.  Covers the fused
peer-memory exchange (csrc/exchange.cuh) and the NCCL all-gather + merge path.  Needs >= 2 GPUs
(`gpurun --gpus 2`); on a single-GPU box there is nothing to shard over and the test is skipped."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "_nccl_worker.py")


def _n_gpus():
    from mlvectordb_b200 import _capi
    return _capi.lib().mlv_device_count()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_search_equals_unsharded(world):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    port = _free_port()
    procs = [subprocess.Popen([sys.executable, WORKER, str(r), str(world), str(port)], stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True, cwd=ROOT) for r in range(world)]
    for r, p in enumerate(procs):
        try:
            out, err = p.communicate(timeout=420)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        assert p.returncode == 0, f"rank {r}:\n{err[-3000:]}"
        assert f"rank {r} ok" in out


def test_sharded_protocol_index_on_one_rank_equals_gpu_index():
    """``ShardedGpuIndex`` with a world of one (no process group): the reference's index / query-processor cases and a
    call-for-call comparison with ``GpuIndex`` -- the protocol host logic on the real device path, visible on a 1-GPU box."""
    import numpy as np
    import torch
    import test_gpu_dropin as TD
    from _refshim import QueryProcessor, Storage, Vector
    from mlvectordb_b200 import GpuIndex, VectorDTO
    from mlvectordb_b200.sharded_index import ShardedGpuIndex
    from oracle import synthetic

    dev = torch.device("cuda", 0)
    make = lambda space, **kw: ShardedGpuIndex(space=space, device=dev, **kw)   # noqa: E731
    TD.test_find_similar_correctness(QueryProcessor(Storage(), make("cosine")))
    TD.test_namespace_isolation(QueryProcessor(Storage(), make("cosine")))
    TD.test_search_with_many_vectors(QueryProcessor(Storage(), make("cosine")))
    TD.test_search_with_few_vectors(QueryProcessor(Storage(), make("cosine")))
    n, dim = 5000, 48
    X = synthetic.rows(17, 0, n, dim, scaled=True)
    vecs = [Vector(values=X[i].tolist(), metadata={"b": i % 10}) for i in range(n)]
    a, b = make("cosine"), GpuIndex(space="cosine")
    for lo, hi in ((0, 3), (3, 2000), (2000, n)):
        a.add(vecs[lo:hi], "ns")
        b.add(vecs[lo:hi], "ns")
    Q = synthetic.queries(17, 4, dim)

    def same(k, **kw):
        for q in Q:
            x = a.search(VectorDTO(values=q), k, "ns", "cosine", **kw)
            y = b.search(VectorDTO(values=q), k, "ns", "cosine", **kw)
            assert [h.vector_id for h in x] == [h.vector_id for h in y] and [h.score for h in x] == [h.score for h in y]

    same(10)
    same(100)
    same(5, filter={"b": 3})
    a.remove([v.id for v in vecs[::4]], "ns")        # 25 % >= 0.2: both compact
    b.remove([v.id for v in vecs[::4]], "ns")
    assert a.info("ns")["tombstones"] == 0 and a.info("ns")["live"] == n - len(vecs[::4])
    same(10)
    x = a.search_async(VectorDTO(values=Q[0]), 10, "ns", "cosine").result()
    assert [h.vector_id for h in x] == [h.vector_id for h in b.search(VectorDTO(values=Q[0]), 10, "ns", "cosine")]
    radius = 1.0 - b.search(VectorDTO(values=Q[1]), 20, "ns", "cosine")[-1].score
    ra, rb = a.range_search(VectorDTO(values=Q[1]), radius, "ns", "cosine"), b.range_search(VectorDTO(values=Q[1]), radius, "ns", "cosine")
    assert [h.vector_id for h in ra] == [h.vector_id for h in rb] and len(ra) >= 20
    a.close()
    b.close()
