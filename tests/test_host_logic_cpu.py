"""CPU: the HOST logic of the drop-in -- ``GpuIndex`` (id maps, counters, thresholds, compaction bookkeeping, metadata
codec, filter cache), ``GpuQueryProcessor``, ``GpuRestAPI`` and snapshots -- with the device shard replaced by an
oracle-backed stand-in (``tests/_fake_shard.py``).  The same test bodies run against the real CUDA library in the
``-m gpu`` suite (``test_gpu_dropin.py``, ``test_gpu_query_processor.py``, ``test_gpu_rest.py``, ``test_gpu_columns.py``);
here they keep the Python around the C ABI honest on a box without a GPU, including the replay of what the UNMODIFIED
reference wrappers returned (``tests/golden/reference_wrappers.json``)."""
import numpy as np
import pytest

import _fake_shard
import test_gpu_columns as TC
import test_gpu_dropin as TD
import test_gpu_query_processor as TQ
from _refshim import InMemoryStorage, QueryProcessor, Storage


@pytest.fixture(autouse=True)
def fake_device(monkeypatch):
    import mlvectordb_b200.index as index_module
    monkeypatch.setattr(index_module, "DeviceShard", _fake_shard.FakeShard)
    yield


@pytest.mark.parametrize("case", TD._golden()["cases"], ids=lambda c: c["name"])
def test_golden_reference_wrapper_replay_host_logic(case):
    TD.test_golden_reference_wrapper_replay(case)


def test_reference_index_and_processor_cases():
    TD.test_golden_query_processor()
    for auto_compact in (True, False):
        TD.test_delete_removes_from_storage_and_index(auto_compact)
    for make in (lambda: QueryProcessor(Storage(), TD._index("cosine")),):
        TD.test_find_similar_correctness(make())
        TD.test_namespace_isolation(make())
        TD.test_search_with_many_vectors(make())
        TD.test_search_with_few_vectors(make())


def test_additive_surface_and_async():
    TD.test_additive_surface_batch_filter_range_dimension()
    TD.test_search_async_equals_search()


def test_metadata_filters_through_index_and_processor():
    TC.test_index_filters_by_metadata_on_the_device("cosine")
    from oracle import synthetic
    n, dim = 6000, 32                                   # the `loaded` fixture of test_gpu_query_processor.py
    X = synthetic.rows(61, 0, n, dim, scaled=True)
    md = [{"bucket": int(i % 10), "parity": "even" if i % 2 == 0 else "odd"} for i in range(n)]
    p = TQ._processor("cosine")
    loaded = (p, X, md, p.upsert_matrix(X, "bulk", metadata=md))
    TQ.test_bulk_ingest_and_plain_search(loaded)
    TQ.test_metadata_filter_dict_and_predicate(loaded)
    TQ.test_batch_equals_single_and_ids_only(loaded)


def test_snapshot_round_trip_host_side(tmp_path):
    from uuid import UUID
    from mlvectordb_b200 import GpuIndex, VectorDTO
    from oracle import synthetic
    n, dim, k = 3000, 24, 5
    X = synthetic.rows(31, 0, n, dim, scaled=True)
    buckets = synthetic.buckets(31, 0, n).astype(np.int64)
    names = np.array(["a", "b", "c"])[np.arange(n) % 3]
    idx = GpuIndex(space="cosine", auto_compact=False)
    ids = idx.add_matrix(X, "big", columns={"bucket": buckets, "name": names})
    idx.add([_V(X[i, :8], {"tag": ("t", i % 2)}) for i in range(50)], "small")
    gone = [UUID(bytes=ids[i].tobytes()) for i in range(0, n, 7)]
    idx.remove(gone, "big")
    Q = synthetic.queries(31, 3, dim)
    cons = {"bucket": ("<", 10), "name": "b"}
    before = idx.search_batch(Q, k, "big", filter=cons)
    before_plain = idx.search_batch(Q, k, "big")
    manifest = idx.save(str(tmp_path / "snap"))
    assert [m["name"] for m in manifest["namespaces"]] == ["big", "small"]
    back = GpuIndex.load(str(tmp_path / "snap"))
    assert back._space == "cosine" and back._auto_compact is False and sorted(back.namespaces()) == ["big", "small"]
    assert back.info("big")["rows"] == n and back.info("big")["tombstones"] == len(gone)
    for a, b in zip(back.search_batch(Q, k, "big"), before_plain):
        assert np.array_equal(a, b)
    for a, b in zip(back.search_batch(Q, k, "big", filter=cons), before):
        assert np.array_equal(a, b)
    assert sorted(back.metadata_columns("big")) == ["bucket", "name"] and back.metadata_columns("small") == ["tag"]
    hit = back.search(VectorDTO(values=X[3, :8]), 1, "small", "cosine", filter={"tag": ("t", 1)})
    assert len(hit) == 1
    victim = back.search(VectorDTO(values=Q[0]), k, "big", "cosine")[0].vector_id
    back.remove([victim], "big")                     # the id map is rebuilt lazily from the restored id table
    assert victim not in {r.vector_id for r in back.search(VectorDTO(values=Q[0]), k, "big", "cosine")}
    assert back.is_rebuild_required("big") == idx.is_rebuild_required("big") or back.is_rebuild_required("big")
    # saving over the snapshot writes a new generation beside the old one and switches the manifest last:
    # a save that dies before the switch leaves the previous snapshot loadable, and stale files are cleaned up
    import os
    snap = str(tmp_path / "snap")
    files0 = sorted(os.listdir(snap))
    assert all(f.startswith("g0.") or f == "manifest.json" for f in files0)
    import mlvectordb_b200.snapshot as snapshot
    real_replace = os.replace
    def boom(*a, **kw):
        raise OSError("crash before the manifest switch")
    snapshot.os.replace = boom
    try:
        with pytest.raises(OSError):
            back.save(snap)
    finally:
        snapshot.os.replace = real_replace
    again = GpuIndex.load(snap)                       # still the first snapshot (victim not yet removed in it)
    assert again.info("big")["tombstones"] == len(gone)
    m2 = back.save(snap)
    assert m2["generation"] == 1 and all(f.startswith("g1.") or f == "manifest.json" for f in os.listdir(snap))
    assert GpuIndex.load(snap).info("big")["tombstones"] == len(gone) + 1


def test_http_surface_over_the_real_processor():
    from fastapi.testclient import TestClient
    from mlvectordb_b200 import GpuIndex, GpuQueryProcessor
    from mlvectordb_b200.rest_api import GpuRestAPI
    import test_gpu_rest as TR
    qp = GpuQueryProcessor(InMemoryStorage(), GpuIndex(space="cosine"))
    client = TestClient(GpuRestAPI(qp, log_level="WARNING").get_app())
    TR.test_ingest_search_filter_range_delete_over_http((client, qp))


class _V:
    def __init__(self, values, metadata=None):
        import uuid
        self.id, self.values, self.metadata = uuid.uuid4(), np.asarray(values, np.float32), metadata


# ---------------------------------------------------------------------------- MultiGpuIndex (one process, several GPUs)
def _numpy_fanout(holders, queries, k, filters):
    """stands in for MultiGpuIndex._device_fanout: per-part search, candidates merged by (distance, part, row)"""
    from mlvectordb_b200.multi import PART_SHIFT
    nq = queries.shape[0]
    D = np.full((nq, k), np.inf, np.float32)
    R = np.full((nq, k), -1, np.int64)
    C = np.zeros(nq, np.int32)
    per = [ns.shard.search(queries, k, flt) for (_, _, ns), flt in zip(holders, filters)]
    for q in range(nq):
        d = np.concatenate([p[0][q, :p[2][q]] for p in per])
        r = np.concatenate([p[1][q, :p[2][q]] + (h[0] << PART_SHIFT) for p, h in zip(per, holders)])
        order = np.lexsort((r, d))[:k]
        D[q, :len(order)], R[q, :len(order)], C[q] = d[order], r[order], len(order)
    return D, R, C


def _numpy_order_pairs(holder, dists, rows):
    order = np.lexsort((np.arange(len(rows)), dists))
    return dists[order], rows[order]


def _multi(space, n_parts=3, **kw):
    from mlvectordb_b200 import MultiGpuIndex
    return MultiGpuIndex(space=space, devices=list(range(n_parts)), fanout=_numpy_fanout, order_pairs=_numpy_order_pairs, **kw)


def test_multi_gpu_index_equals_single_index_behind_the_reference_calls(monkeypatch):
    """Same adds / removes / searches through one GpuIndex and through a 3-part MultiGpuIndex: same ids and scores."""
    from mlvectordb_b200 import GpuIndex, VectorDTO
    from oracle import synthetic
    rng = np.random.default_rng(4)
    for space in ("cosine", "l2"):
        one, many = GpuIndex(space=space), _multi(space)
        X = synthetic.rows(9, 0, 900, 20, scaled=True)
        vecs = [_V(X[i], {"bucket": i % 7, "color": ["r", "g", "b"][i % 3]}) for i in range(900)]
        for lo, hi in ((0, 1), (1, 4), (4, 400), (400, 900)):          # single rows, small and large blocks
            one.add(vecs[lo:hi], "ns")
            many.add(vecs[lo:hi], "ns")
        per_part = many.info("ns")["rows_per_device"]
        assert sum(per_part) == 900 and max(per_part) - min(per_part) <= 1
        Q = synthetic.queries(9, 5, 20)
        def same(k, flt=None, metric=space):
            for q in Q:
                a = one.search(VectorDTO(values=q), k, "ns", metric, filter=flt)
                b = many.search(VectorDTO(values=q), k, "ns", metric, filter=flt)
                assert [r.vector_id for r in a] == [r.vector_id for r in b]
                assert [r.score for r in a] == pytest.approx([r.score for r in b], rel=1e-6, abs=1e-7)
        same(10)
        same(1000)                                                       # clamped to the live count
        same(5, {"bucket": 3, "color": "g"})
        same(5, {"bucket": ("<", 2)})
        allowed = {v.id for v in vecs[::11]}
        same(5, lambda u: u in allowed)
        gone = [vecs[i].id for i in rng.choice(900, 150, replace=False)]
        one.remove(gone, "ns")
        many.remove(gone + [vecs[0].id.__class__(int=5)], "ns")         # unknown id: ignored
        assert many.info("ns")["live"] == 750
        same(10)
        tenth = one.search(VectorDTO(values=Q[0]), 10, "ns", space)[-1].score
        radius = (1 - tenth if space == "cosine" else tenth) * (1 + 1e-6) + 1e-7
        hits_one = one.range_search(VectorDTO(values=Q[0]), radius, "ns", space)
        hits_many = many.range_search(VectorDTO(values=Q[0]), radius, "ns", space)
        assert [h.vector_id for h in hits_one] == [h.vector_id for h in hits_many] and len(hits_one) >= 10
        rows, scores, counts = many.search_batch(Q, 4, "ns")
        assert [u for u in many.uuids_of("ns", rows[2])] == [r.vector_id for r in one.search(VectorDTO(values=Q[2]), 4, "ns", space)]
        assert many.search(VectorDTO(values=[1.0, 2.0]), 3, "ns", space) == [] and many.search(VectorDTO(values=Q[0]), 3, "nope", space) == []
        assert many.dimension("ns") == 20 and many.dimension("nope") is None and sorted(many.metadata_columns("ns")) == ["bucket", "color"]
        many.rebuild({"a": vecs[:10], "b": vecs[10:13]}, metric=space)
        assert sorted(many.namespaces()) == ["a", "b"] and many.info("a")["rows"] == 10 and many.info("b")["rows_per_device"] == [1, 1, 1]
        assert many.search(VectorDTO(values=X[11]), 1, "b", space)[0].vector_id == vecs[11].id
        assert many.is_rebuild_required("a") is False
        one.close()
        many.close()


def test_multi_gpu_index_behind_the_query_processors():
    """The reference-shaped ``QueryProcessor`` and ``GpuQueryProcessor`` over a MultiGpuIndex: the reference's own
    query-processor cases, plus a metadata filter decided per part."""
    from mlvectordb_b200 import GpuQueryProcessor, VectorDTO
    TD.test_find_similar_correctness(QueryProcessor(Storage(), _multi("cosine")))
    TD.test_namespace_isolation(QueryProcessor(Storage(), _multi("cosine")))
    TD.test_search_with_many_vectors(QueryProcessor(Storage(), _multi("cosine")))
    qp = GpuQueryProcessor(InMemoryStorage(), _multi("cosine", n_parts=2))
    qp.upsert_many([VectorDTO(values=[1, 0, 0], metadata={"t": "x"}), VectorDTO(values=[0.9, 0.1, 0], metadata={"t": "y"}),
                    VectorDTO(values=[0, 1, 0], metadata={"t": "x"}), VectorDTO(values=[0, 0, 1], metadata={"t": ["unhashable"]})])
    hits = qp.find_similar(VectorDTO(values=[1, 0, 0]), 3, filter={"t": "x"})
    assert [h["metadata"]["t"] for h in hits] == ["x", "x"] and hits[0]["score"] == pytest.approx(1.0)
    hits = qp.find_similar(VectorDTO(values=[1, 0, 0]), 3, filter=lambda md: md["t"] == "y")     # host predicate path
    assert [h["metadata"]["t"] for h in hits] == ["y"]
    deleted = qp.delete([hits[0]["id"]])
    assert deleted == [hits[0]["id"]] and len(qp.find_similar(VectorDTO(values=[1, 0, 0]), 10)) == 3


def test_multi_gpu_index_bulk_ingest_and_dimension_guard():
    """``GpuQueryProcessor.upsert_matrix`` over a ``MultiGpuIndex`` (metadata cut per part), and a
    wrong-dimension block is refused even by a part that does not hold the namespace yet -- before the storage sees it."""
    from mlvectordb_b200 import GpuQueryProcessor, VectorDTO
    from oracle import synthetic
    many = _multi("cosine", n_parts=3)
    storage = InMemoryStorage()
    qp = GpuQueryProcessor(storage, many)
    X = synthetic.rows(2, 0, 50, 8, scaled=True)
    ids = qp.upsert_matrix(X, "bulk", metadata=[{"b": i % 4} for i in range(50)])
    assert len(ids) == 50 and many.info("bulk")["rows"] == 50 and many.metadata_columns("bulk") == ["b"]
    hits = qp.find_similar(VectorDTO(values=X[7].tolist()), 3, namespace="bulk", filter={"b": 3})
    assert hits[0]["id"] == ids[7] and all(h["metadata"]["b"] == 3 for h in hits)
    stored_before = qp.get_namespace_count("bulk")
    with pytest.raises(RuntimeError, match="dimensionality"):
        qp.upsert_matrix(synthetic.rows(2, 0, 2, 9), "bulk")           # 2 rows -> only two parts get a slice
    assert qp.get_namespace_count("bulk") == stored_before and many.info("bulk")["rows"] == 50
    one_row = _multi("l2", n_parts=3)
    one_row.add([_V(np.ones(4))], "ns")                                  # lives on part 0 only
    with pytest.raises(RuntimeError, match="dimensionality"):
        one_row.add([_V(np.ones(5))], "ns")                              # would have landed on the empty part 1
    with pytest.raises(RuntimeError, match="dimensionality"):
        one_row.add([_V(np.ones(4)), _V(np.ones(5))], "fresh")
    assert one_row.search(VectorDTO(values=[1.0] * 5), 1, "ns", "l2") == []
    many.close()
    one_row.close()


def test_http_surface_over_a_multi_gpu_index():
    """``GpuRestAPI`` -> ``GpuQueryProcessor`` -> ``MultiGpuIndex``: the same HTTP scenario as over one ``GpuIndex``."""
    from fastapi.testclient import TestClient
    from mlvectordb_b200 import GpuQueryProcessor
    from mlvectordb_b200.rest_api import GpuRestAPI
    import test_gpu_rest as TR
    qp = GpuQueryProcessor(InMemoryStorage(), _multi("cosine", n_parts=2))
    client = TestClient(GpuRestAPI(qp, log_level="WARNING").get_app())
    TR.test_ingest_search_filter_range_delete_over_http((client, qp))
