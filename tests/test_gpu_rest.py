"""GPU: the HTTP surface end to end -- ``GpuRestAPI`` over ``GpuQueryProcessor`` over ``GpuIndex``
(SURVEY.md section 8f rank 1).  What comes back over HTTP equals what the query processor returns
directly and agrees with the oracle."""
import numpy as np
import pytest

from oracle import exact, synthetic
from _refshim import InMemoryStorage

pytestmark = pytest.mark.gpu


@pytest.fixture()
def served():
    from fastapi.testclient import TestClient
    from mlvectordb_b200 import GpuIndex, GpuQueryProcessor
    from mlvectordb_b200.rest_api import GpuRestAPI
    qp = GpuQueryProcessor(InMemoryStorage(), GpuIndex(space="cosine"))
    client = TestClient(GpuRestAPI(qp, log_level="WARNING").get_app())
    yield client, qp
    qp._index.close()


def test_ingest_search_filter_range_delete_over_http(served):
    client, qp = served
    n, dim, k = 400, 24, 5
    X = synthetic.rows(5, 0, n, dim, scaled=True)
    colors = ["red", "green", "blue"]
    vectors = [{"values": X[i].tolist(), "metadata": {"color": colors[i % 3], "bucket": i % 10}} for i in range(n)]
    assert client.put("/vectors/batch", json={"vectors": vectors[:300]}).status_code == 200
    for v in vectors[300:303]:
        assert client.post("/vectors", json=v).status_code == 201
    assert client.put("/vectors/batch", json={"vectors": vectors[303:]}).status_code == 200
    q = synthetic.queries(5, 3, dim)
    # plain top-k: the reference request, the reference response shape, oracle ids and scores
    body = client.post("/search", json={"query": q[0].tolist(), "top_k": k, "metric": "cosine"}).json()
    L, D = exact.knn(X, q[:1], k, "cosine")
    got_rows = [int(np.flatnonzero((X == np.asarray(h["values"], np.float32)).all(axis=1))[0]) for h in body]
    assert exact.check_topk_parity(got_rows, [1 - h["score"] for h in body], L[0], D[0]) is None
    assert all(h["metadata"] == vectors[r]["metadata"] for h, r in zip(body, got_rows))
    # filter: evaluated on the device columns; ids only
    flt = {"color": "red", "bucket": ["<", 5]}
    mask = np.array([(i % 3 == 0) and (i % 10 < 5) for i in range(n)])
    body = client.post("/search", json={"query": q[1].tolist(), "top_k": k, "filter": flt, "include_values": False}).json()
    direct = qp.find_similar(type("Q", (), {"values": q[1], "metadata": {}})(), k, "default", "cosine",
                             filter={"color": "red", "bucket": ("<", 5)}, enrich=False)
    assert [h["id"] for h in body] == [str(h["id"]) for h in direct] and all(h["values"] == [] for h in body)
    L, D = exact.knn(X, q[1:2], k, "cosine", allow=mask)
    assert np.allclose([1 - h["score"] for h in body], D[0], rtol=1e-5, atol=1e-6)
    assert qp._index.where("default", {"color": "red"}) is not None          # the device path, not the host fallback
    # range / similarity
    radius = float(D[0][-1]) + 0.05
    hits = client.post("/query/range", json={"vector": q[1].tolist(), "radius": radius, "metric": "cosine"}).json()
    want = exact.range_search(X, q[1:2], radius, "cosine")[0]
    assert hits["count"] == len(want[0])
    sim = client.post("/query/similarity", json={"vector": q[1].tolist(), "threshold": 1 - radius, "include_values": False}).json()
    assert sim["count"] == hits["count"] and [h["id"] for h in sim["results"]] == [h["id"] for h in hits["results"]]
    assert all(h["score"] >= 1 - radius - 1e-6 for h in sim["results"])
    # batch
    many = client.post("/search/batch", json={"queries": q.tolist(), "top_k": k}).json()
    L, D = exact.knn(X, q, k, "cosine")
    for i in range(3):
        assert np.allclose([1 - h["score"] for h in many[i]], D[i], rtol=1e-5, atol=1e-6)
    # delete, then the hit is gone; statistics reports the device state
    victim = many[0][0]["id"]
    assert client.request("DELETE", "/vectors", json={"ids": [victim]}).json()["message"] == "1 vectors deleted"
    again = client.post("/search", json={"query": q[0].tolist(), "top_k": k, "include_values": False}).json()
    assert victim not in [h["id"] for h in again]
    stats = client.get("/statistics").json()["namespaces"]["default"]
    assert stats["rows"] == n and stats["live"] == n - 1 and sorted(stats["metadata_columns"]) == ["bucket", "color"]
    assert client.get("/namespaces").json() == {"namespaces": ["default"]}
    # wrong dimension: the reference answers [] (index.py:110-119), not an error
    assert client.post("/search", json={"query": [1.0, 2.0], "top_k": 3}).json() == []
