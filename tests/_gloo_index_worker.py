"""One rank of the world_size-2 gloo test of ``ShardedGpuIndex`` (launched by tests/test_sharded_gloo.py).

usage: _gloo_index_worker.py RANK WORLD PORT
The device side is played by the oracle-backed ``FakeShard`` (tests/_fake_shard.py) and a numpy restatement of the merge
kernel's contract; the SPMD host logic -- block splitting over the ranks, replicated id tables, tombstone mirrors,
thresholds, compaction renumbering, rebuild, global-row decoding -- is the product's, and is compared call for call
with a single-process ``GpuIndex`` over the same stand-in.
"""
import os
import sys
from uuid import UUID

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class V:
    def __init__(self, uid, values, metadata=None):
        self.id, self.values, self.metadata = uid, values, metadata or {}


def main():
    rank, world, port = (int(x) for x in sys.argv[1:4])
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import _fake_shard
    import mlvectordb_b200.index as index_module
    from _gloo_worker import numpy_merge
    from mlvectordb_b200 import GpuIndex, VectorDTO
    from mlvectordb_b200.multi import PART_SHIFT
    from mlvectordb_b200.sharded import ShardedIndex
    from mlvectordb_b200.sharded_index import ShardedGpuIndex
    from oracle import synthetic

    index_module.DeviceShard = _fake_shard.FakeShard          # the single-process comparison index

    def factory(dim, space, capacity):
        fs = _fake_shard.FakeShard(dim, space, row_base=rank << PART_SHIFT)

        def local_search(q, k):
            d, r, c = fs.search(q.numpy(), k)
            return torch.from_numpy(d), torch.from_numpy(r), torch.from_numpy(c)

        def local_range(q, radius):
            return fs.range_search(q, radius) if fs.live else [(np.empty(0, np.float32), np.empty(0, np.int64)) for _ in range(len(q))]

        def order_hits(d, r):
            dn, rn = d.numpy(), r.numpy()
            order = np.lexsort((rn, dn, rn < 0))
            return torch.from_numpy(dn[order]), torch.from_numpy(rn[order])

        return ShardedIndex(dim, space, 0, device=None, local_search=local_search, merge=numpy_merge, local_range=local_range,
                            order_hits=order_hits, row_bases=[r << PART_SHIFT for r in range(world)], shard=fs)

    dim, space = 12, "cosine"
    X = synthetic.rows(3, 0, 400, dim, scaled=True)
    ids = [UUID(int=1000 + i) for i in range(400)]
    sharded = ShardedGpuIndex(space=space, rebuild_threshold=0.2, searcher_factory=factory)
    single = GpuIndex(space=space, rebuild_threshold=0.2)
    assert sharded.world == world and sharded._space == space

    def both(fn):
        return fn(sharded), fn(single)

    def same_hits(a, b):
        assert [h.vector_id for h in a] == [h.vector_id for h in b], ([h.vector_id.int for h in a], [h.vector_id.int for h in b])
        assert np.allclose([h.score for h in a], [h.score for h in b], rtol=1e-6, atol=1e-7)

    q = VectorDTO(values=synthetic.queries(3, 1, dim)[0].tolist())
    assert sharded.search(q, 5, "a", space) == [] and sharded.dimension("a") is None
    for lo, hi in ((0, 7), (7, 150), (150, 151), (151, 400)):       # uneven blocks: water filling keeps the ranks level
        both(lambda ix: ix.add([V(ids[i], X[i], {"b": i % 5}) for i in range(lo, hi)], "a"))
    counts = sharded.info("a")["rows_per_rank"]
    assert sum(counts) == 400 and max(counts) - min(counts) <= 1, counts
    for k in (1, 10, 500):
        same_hits(*both(lambda ix: ix.search(q, k, "a", space)))
    assert len(sharded.search(q, 500, "a", space)) == 400                              # top_k clamped to the live count
    assert sharded.search(VectorDTO(values=[0.0] * (dim + 1)), 3, "a", space) == []  # wrong dimension -> []
    same_hits(sharded.search_async(q, 10, "a", space).result(), single.search(q, 10, "a", space))
    # exact stored row first, l2-style metric toggle untouched
    hit = sharded.search(VectorDTO(values=X[123].tolist()), 1, "a", space)[0]
    assert hit.vector_id == ids[123] and abs(hit.score - 1.0) < 1e-5
    # removal: unknown ids ignored, removed ids never return, threshold -> per-rank compaction with renumbering
    both(lambda ix: ix.remove([ids[5], ids[123], UUID(int=7)], "a"))
    assert ids[123] not in [h.vector_id for h in sharded.search(VectorDTO(values=X[123].tolist()), 400, "a", space)]
    both(lambda ix: ix.remove(ids[200:300], "a"))                 # 102 / 400 >= 0.2: both compact
    inf = sharded.info("a")
    assert inf["tombstones"] == 0 and inf["live"] == 298 and sum(inf["rows_per_rank"]) == 298
    assert not sharded.is_rebuild_required("a")
    same_hits(*both(lambda ix: ix.search(q, 25, "a", space)))
    both(lambda ix: ix.remove(ids[300:310], "a"))                 # lookup table rebuilt lazily after the compaction
    same_hits(*both(lambda ix: ix.search(q, 25, "a", space)))
    # range search: every rank's hits, ordered
    radius = 1.0 - single.search(q, 12, "a", space)[-1].score
    same_hits(*both(lambda ix: ix.range_search(q, radius, "a", space)))
    # namespaces are independent; bulk ingest; rebuild replaces everything with the new metric
    Y = synthetic.rows(4, 0, 60, dim)
    got = sharded.add_matrix(Y, "b", ids=[UUID(int=5000 + i) for i in range(60)])
    assert got.shape == (60, 16) and sharded.dimension("b") == dim and sorted(sharded.namespaces()) == ["a", "b"]
    rows, scores, cnt = sharded.search_batch(Y[:3], 2, "b")
    assert [u.int for u in sharded.uuids_of("b", rows[:, 0])] == [5000, 5001, 5002] and (cnt == 2).all()
    try:
        sharded.add([V(UUID(int=1), np.zeros(dim + 1, np.float32))], "a")
        raise AssertionError("wrong dimension accepted")
    except RuntimeError as e:
        assert "dimensionality" in str(e)
    src = {"c": [V(ids[i], X[i]) for i in range(50)]}
    both(lambda ix: ix.rebuild(src, "l2"))
    assert sharded.namespaces() == ["c"]
    same_hits(*both(lambda ix: ix.search(q, 5, "c", "l2")))
    no_compact = ShardedGpuIndex(space="l2", auto_compact=False, searcher_factory=factory)
    no_compact.add([V(ids[i], X[i]) for i in range(10)], "z")
    no_compact.remove(ids[:2], "z")
    assert no_compact.is_rebuild_required("z")                    # reference index.py:86-89 flag behaviour
    # ranks agree on everything
    mine = torch.tensor([h.vector_id.int % (1 << 62) for h in sharded.search(q, 5, "c", "l2")], dtype=torch.int64)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    assert all(torch.equal(g, mine) for g in gathered)
    sharded.close()
    dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
