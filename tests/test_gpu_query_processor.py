"""SURVEY.md section 8f (ranks 1-3) on a GPU: ``GpuQueryProcessor`` keeps the reference ``QueryProcessor``'s
behaviour (reference ``tests/test_query_processor.py`` restated) and its additive entry points --
metadata filter, batches, range, bulk ingest, ids-only responses -- agree with the CPU oracle."""
import numpy as np
import pytest

from oracle import exact, synthetic
from _refshim import Storage

pytestmark = pytest.mark.gpu


class _Storage(Storage):
    """the reference storage calls GpuQueryProcessor adds on top of tests/_refshim.Storage"""

    @property
    def list_namespaces(self):
        return list(self._data)

    def get_storage_info(self):
        return {"total_vectors": sum(len(d) for d in self._data.values())}


def _processor(space="cosine", **kw):
    from mlvectordb_b200 import GpuIndex, GpuQueryProcessor
    return GpuQueryProcessor(_Storage(), GpuIndex(space=space, **kw))


def _dto(values, metadata=None):
    from mlvectordb_b200 import VectorDTO
    return VectorDTO(values=values, metadata=metadata or {})


# ---- reference tests/test_query_processor.py, restated ------------------------------------------
def test_find_similar_correctness():
    p = _processor()
    p.upsert_many([_dto([1, 0, 0], {"n": "a"}), _dto([0.9, 0.1, 0], {"n": "b"}), _dto([0, 1, 0], {"n": "c"})])
    res = p.find_similar(_dto([1, 0, 0]), top_k=3)
    assert [r["metadata"]["n"] for r in res] == ["a", "b", "c"]
    assert res[0]["score"] == pytest.approx(1.0, rel=1e-4)
    assert res[0]["score"] >= res[1]["score"] >= res[2]["score"]
    assert set(res[0]) == {"id", "values", "metadata", "score"}


def test_namespace_isolation_and_delete():
    p = _processor()
    p.insert(_dto([1, 0], {"ns": 1}), namespace="one")
    p.insert(_dto([1, 0], {"ns": 2}), namespace="two")
    assert [r["metadata"]["ns"] for r in p.find_similar(_dto([1, 0]), 5, namespace="one")] == [1]
    assert sorted(p.list_namespaces()) == ["one", "two"]
    vid = p.find_similar(_dto([1, 0]), 1, namespace="two")[0]["id"]
    assert list(p.delete([vid], namespace="two")) == [vid]
    assert p.find_similar(_dto([1, 0]), 5, namespace="two") == []
    assert p.get_namespace_count("two") == 0 and p.get_namespace_count("one") == 1
    assert len(p.find_similar(_dto([1, 0]), 5, namespace="one")) == 1      # the other namespace survives (Q6)
    assert p.find_similar(_dto([1, 0]), 5, namespace="missing") == []


# ---- additive entry points vs the oracle ---------------------------------------------------------
@pytest.fixture(scope="module")
def loaded():
    n, dim = 6000, 32
    X = synthetic.rows(61, 0, n, dim, scaled=True)
    md = [{"bucket": int(i % 10), "parity": "even" if i % 2 == 0 else "odd"} for i in range(n)]
    p = _processor("cosine")
    ids = p.upsert_matrix(X, "bulk", metadata=md)
    return p, X, md, ids


def _expect(X, Q, k, allow=None):
    L, D = exact.knn(X, Q, k, "cosine", allow=allow)
    return L, [1.0 - np.asarray(d, np.float64) for d in D]


def test_bulk_ingest_and_plain_search(loaded):
    p, X, md, ids = loaded
    assert p.get_namespace_count("bulk") == len(X) and len(set(ids)) == len(X)
    Q = synthetic.queries(62, 3, X.shape[1])
    L, S = _expect(X, Q, 10)
    for i in range(3):
        res = p.find_similar(_dto(Q[i].tolist()), 10, namespace="bulk", metric="cosine")
        assert [r["id"] for r in res] == [ids[j] for j in L[i]]
        assert np.allclose([r["score"] for r in res], S[i], rtol=1e-5, atol=1e-6)
        assert all(np.array_equal(r["values"], X[j]) and r["metadata"] == md[j] for r, j in zip(res, L[i]))


def test_metadata_filter_dict_and_predicate(loaded):
    p, X, md, ids = loaded
    Q = synthetic.queries(63, 2, X.shape[1])
    allow = np.array([m["bucket"] == 3 and m["parity"] == "odd" for m in md])
    L, S = _expect(X, Q, 10, allow)
    for flt in ({"bucket": 3, "parity": "odd"}, lambda m: m["bucket"] == 3 and m["parity"] == "odd"):
        for i in range(2):
            res = p.find_similar(_dto(Q[i].tolist()), 10, namespace="bulk", metric="cosine", filter=flt)
            assert [r["id"] for r in res] == [ids[j] for j in L[i]]
            assert all(r["metadata"]["bucket"] == 3 for r in res)
    assert p.find_similar(_dto(Q[0].tolist()), 10, namespace="bulk", filter={"bucket": 99}) == []
    # dict filters are decided on the device columns, prepared once and reused until the namespace changes
    assert len(p._index._ns["bulk"].where_cache) == 2 and not p._filters
    # ... and on the host when the device cannot decide them (None also matches rows without the key)
    assert p.find_similar(_dto(Q[0].tolist()), 10, namespace="bulk", filter={"nokey": None, "bucket": 3}, enrich=False)
    assert sum(1 for k in p._filters if k[0] == "bulk") == 1


def test_batch_equals_single_and_ids_only(loaded):
    p, X, md, ids = loaded
    Q = synthetic.queries(64, 40, X.shape[1])          # >= 32 queries: tensor-core path when the shape allows
    batch = p.find_similar_batch([_dto(q.tolist()) for q in Q], 5, namespace="bulk", metric="cosine", enrich=False)
    assert len(batch) == 40 and all(set(r) == {"id", "score"} for hits in batch for r in hits)
    for i in (0, 7, 39):
        single = p.find_similar(_dto(Q[i].tolist()), 5, namespace="bulk", metric="cosine")
        assert [r["id"] for r in batch[i]] == [r["id"] for r in single]
        assert [r["score"] for r in batch[i]] == pytest.approx([r["score"] for r in single], rel=1e-6)
    filt = p.find_similar_batch(Q[:3], 5, namespace="bulk", filter={"parity": "even"})
    assert all(r["metadata"]["parity"] == "even" for hits in filt for r in hits)


def test_range_query(loaded):
    p, X, md, ids = loaded
    q = synthetic.queries(65, 1, X.shape[1])
    L, S = _expect(X, q, 20)
    radius = float(1.0 - (S[0][11] + S[0][12]) / 2)      # between the 12th and 13th hit
    res = p.find_in_range(_dto(q[0].tolist()), radius, namespace="bulk", metric="cosine")
    assert [r["id"] for r in res] == [ids[j] for j in L[0][:12]]
    assert all(r["score"] >= 1.0 - radius - 1e-6 for r in res)
    odd = p.find_in_range(_dto(q[0].tolist()), radius, namespace="bulk", metric="cosine", filter={"parity": "odd"})
    assert [r["id"] for r in odd] == [ids[j] for j in L[0][:12] if md[j]["parity"] == "odd"]


def test_filters_are_dropped_when_the_namespace_changes():
    p = _processor("l2")
    X = synthetic.rows(66, 0, 500, 16)
    ids = p.upsert_matrix(X, "ns", metadata=[{"g": i % 4} for i in range(500)])
    q = _dto(X[8].tolist())
    assert p.find_similar(q, 3, "ns", "l2", filter={"g": 0})[0]["id"] == ids[8]
    p.delete([ids[8]], "ns")
    assert p._filters == {}
    res = p.find_similar(q, 3, "ns", "l2", filter={"g": 0})
    assert ids[8] not in [r["id"] for r in res] and all(r["metadata"]["g"] == 0 for r in res)
    p.delete(ids[:200], "ns")                            # crosses rebuild_threshold: device compaction
    res = p.find_similar(q, 3, "ns", "l2", filter={"g": 0})
    assert len(res) == 3 and all(r["metadata"]["g"] == 0 for r in res)
