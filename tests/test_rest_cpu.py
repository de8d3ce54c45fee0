"""CPU: ``GpuRestAPI`` keeps the reference's route table, models and messages
(``src/mlvectordb/api/rest_api.py:96-311``) and maps the additive request fields onto the query processor.
A recording stand-in plays the processor, so no GPU is touched here."""
import uuid

import numpy as np
import pytest

fastapi = pytest.importorskip("fastapi")
from fastapi.testclient import TestClient  # noqa: E402

from mlvectordb_b200.rest_api import GpuRestAPI  # noqa: E402


class _Recorder:
    """QueryProcessorProtocol-shaped (reference interfaces/query_processor.py:7-11) plus the additive entry points."""

    def __init__(self):
        self.calls = []
        self.ids = [uuid.uuid4() for _ in range(3)]

    def _hits(self, n, enrich=True):
        out = []
        for i in range(n):
            h = {"id": self.ids[i], "score": 1.0 - 0.1 * i}
            if enrich:
                h.update(values=np.arange(4, dtype=np.float32) + i, metadata={"i": i})
            out.append(h)
        return out

    def insert(self, dto, namespace="default"):
        self.calls.append(("insert", list(dto.values), dict(dto.metadata), namespace))

    def upsert_many(self, dtos, namespace="default"):
        self.calls.append(("upsert_many", len(dtos), namespace))

    def find_similar(self, query, top_k, namespace="default", metric="cosine", **extra):
        self.calls.append(("find_similar", list(query.values), top_k, namespace, metric, extra))
        return self._hits(min(top_k, 3), extra.get("enrich", True))

    def find_similar_batch(self, queries, top_k, namespace="default", metric="cosine", filter=None, enrich=True):  # noqa: A002
        self.calls.append(("find_similar_batch", len(queries), top_k, namespace, metric, filter, enrich))
        return [self._hits(min(top_k, 2), enrich) for _ in queries]

    def find_in_range(self, query, radius, namespace="default", metric="cosine", **extra):
        self.calls.append(("find_in_range", radius, namespace, metric, extra))
        return self._hits(2, extra.get("enrich", True))

    def delete(self, ids, namespace="default"):
        self.calls.append(("delete", list(ids), namespace))
        return [i for i in ids if i in self.ids]

    def list_namespaces(self):
        return ["default", "other"]

    def get_namespace_vectors(self, namespace):
        return [{"id": self.ids[0], "values": np.ones(4, np.float32), "metadata": {}}]

    def get_storage_info(self):
        return {"total_vectors": 3}


@pytest.fixture()
def client_and_qp():
    qp = _Recorder()
    return TestClient(GpuRestAPI(qp, log_level="WARNING").get_app()), qp


def test_reference_route_table_is_kept(client_and_qp):
    client, _ = client_and_qp
    routes = {(m, r.path) for r in client.app.routes if hasattr(r, "methods") for m in r.methods}
    for want in [("POST", "/vectors"), ("PUT", "/vectors/batch"), ("POST", "/search"), ("DELETE", "/vectors"),
                 ("GET", "/namespaces"), ("GET", "/namespaces/vectors"), ("GET", "/storage/info"), ("GET", "/health"),
                 ("POST", "/log/level")]:
        assert want in routes, want
    for extra in [("POST", "/search/batch"), ("POST", "/query/knn"), ("POST", "/query/range"), ("POST", "/query/similarity"),
                  ("POST", "/query/hybrid"), ("GET", "/statistics")]:
        assert extra in routes, extra


def test_reference_requests_get_reference_responses(client_and_qp):
    client, qp = client_and_qp
    r = client.post("/vectors", params={"namespace": "a"}, json={"values": [1, 2, 3], "metadata": {"k": "v"}})
    assert r.status_code == 201 and r.json() == {"status": "success", "message": "Vector inserted"}
    assert qp.calls[-1] == ("insert", [1.0, 2.0, 3.0], {"k": "v"}, "a")
    r = client.put("/vectors/batch", json={"vectors": [{"values": [1, 2]}, {"values": [3, 4], "metadata": {}}]})
    assert r.status_code == 200 and r.json() == {"status": "success", "message": "2 vectors upserted"}
    # the reference's search body (rest_api.py:22-25) reaches find_similar with the reference's exact keywords
    r = client.post("/search", json={"query": [1, 0, 0, 0], "top_k": 2, "metric": "l2"})
    assert r.status_code == 200
    assert qp.calls[-1] == ("find_similar", [1.0, 0.0, 0.0, 0.0], 2, "default", "l2", {})
    body = r.json()
    assert [set(h) for h in body] == [{"id", "values", "metadata", "score"}] * 2
    assert body[0]["id"] == str(qp.ids[0]) and body[1]["values"] == [1.0, 2.0, 3.0, 4.0] and body[0]["score"] == 1.0
    assert client.post("/search", json={"query": [1.0]}).status_code == 200 and qp.calls[-1][2] == 10   # default top_k
    assert client.post("/search", json={"query": [1.0], "top_k": 0}).status_code == 422
    assert client.post("/search", json={"query": [1.0], "top_k": 1001}).status_code == 422
    r = client.request("DELETE", "/vectors", json={"ids": [str(qp.ids[1]), str(uuid.uuid4())]})
    assert r.json() == {"status": "success", "message": "1 vectors deleted"}
    assert client.request("DELETE", "/vectors", json={"ids": []}).status_code == 400
    assert client.request("DELETE", "/vectors", json={"ids": [str(uuid.uuid4())]}).json()["status"] == "error"
    assert client.get("/namespaces").json() == {"namespaces": ["default", "other"]}
    assert client.get("/namespaces/vectors").json()[0]["values"] == [1.0] * 4
    assert client.get("/storage/info").json() == {"total_vectors": 3}
    assert client.get("/health").json() == {"status": "healthy"}
    assert client.post("/log/level", params={"level": "debug"}).json()["message"] == "Log level set to DEBUG"
    assert client.post("/log/level", params={"level": "loud"}).status_code == 400


def test_additive_fields_reach_the_processor(client_and_qp):
    client, qp = client_and_qp
    r = client.post("/search", json={"query": [1, 0], "top_k": 3, "filter": {"color": "red", "bucket": ["<", 5]},
                                     "include_values": False})
    assert r.status_code == 200
    assert qp.calls[-1][-1] == {"filter": {"color": "red", "bucket": ("<", 5)}, "enrich": False}
    assert r.json()[0]["values"] == [] and r.json()[0]["metadata"] == {}
    r = client.post("/search", json={"query": [1, 0], "radius": 0.25, "metric": "cosine"}, params={"namespace": "n"})
    assert qp.calls[-1] == ("find_in_range", 0.25, "n", "cosine", {})
    r = client.post("/search/batch", json={"queries": [[1, 0], [0, 1], [1, 1]], "top_k": 5, "filter": {"a": 1}})
    assert r.status_code == 200 and len(r.json()) == 3 and len(r.json()[0]) == 2
    assert qp.calls[-1] == ("find_similar_batch", 3, 5, "default", "cosine", {"a": 1}, False)
    # the example client's request shapes (examples/api_client.py:26-74)
    r = client.post("/query/knn", json={"type": "knn", "vector": [1, 0], "k": 2})
    assert r.status_code == 200 and r.json()["count"] == 2 and qp.calls[-1][0] == "find_similar"
    r = client.post("/query/range", json={"type": "range", "vector": [1, 0], "radius": 0.5})
    assert r.json()["type"] == "range" and qp.calls[-1][:2] == ("find_in_range", 0.5)
    assert client.post("/query/range", json={"vector": [1, 0]}).status_code == 400
    r = client.post("/query/similarity", json={"type": "similarity", "vector": [1, 0], "threshold": 0.8, "metric": "cosine"})
    assert r.status_code == 200 and abs(qp.calls[-1][1] - 0.2) < 1e-12
    r = client.post("/query/hybrid", json={"vector": [1, 0], "k": 1, "filter": {"color": "red"}, "namespace": "z"})
    assert qp.calls[-1] == ("find_similar", [1.0, 0.0], 1, "z", "cosine", {"filter": {"color": "red"}})
    assert client.post("/query/hybrid", json={"vector": [1, 0], "k": 1}).status_code == 400
    assert client.get("/statistics").json() == {"namespaces": {}}


def test_processor_errors_become_http_500_like_the_reference(client_and_qp):
    client, qp = client_and_qp

    def boom(*a, **k):
        raise RuntimeError("device lost")
    qp.find_similar = boom
    r = client.post("/search", json={"query": [1.0]})
    assert r.status_code == 500 and r.json()["detail"] == "Search failed: device lost"
