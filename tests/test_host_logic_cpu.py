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


def test_http_surface_over_the_real_processor():
    from fastapi.testclient import TestClient
    from mlvectordb_b200 import GpuIndex, GpuQueryProcessor
    from mlvectordb_b200.rest_api import GpuRestAPI
    import test_gpu_rest as TR
    qp = GpuQueryProcessor(InMemoryStorage(), GpuIndex(space="cosine"))
    client = TestClient(GpuRestAPI(qp, log_level="WARNING").get_app())
    TR.test_ingest_search_filter_range_delete_over_http((client, qp))


class _V:
    def __init__(self, values, metadata):
        import uuid
        self.id, self.values, self.metadata = uuid.uuid4(), np.asarray(values, np.float32), metadata
