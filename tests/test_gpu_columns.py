"""GPU parity of the columnar metadata store and device-evaluated predicates
(``mlv_index_set_column`` / ``mlv_filter_create_where``, ``csrc/column_kernels.cuh``).

Integer work, so the bar is bit-exact: the bitmap ``where_kernel`` writes equals ``oracle.exact.where_mask``
word for word, and a search restricted by it returns the same bits as the same search restricted by a
host-built mask (and the oracle's hnswlib-0.8 ``filter=`` semantics)."""
import numpy as np
import pytest

from oracle import exact, synthetic

pytestmark = pytest.mark.gpu

MISSING = exact.COLUMN_MISSING


def _same(a, b):
    return all(np.array_equal(x, y, equal_nan=True) for x, y in zip(a, b))


def _bits(words, n):
    return np.unpackbits(np.ascontiguousarray(words).view(np.uint8), bitorder="little")[:n].astype(bool)


PREDICATE_SETS = [
    [(0, "==", 7)], [(0, "!=", 7)], [(0, "<", 10)], [(0, "<=", 10)], [(0, ">", 90)], [(0, ">=", 90)], [(0, "between", 20, 29)],
    [(0, "between", 5, 4)], [(1, "==", -3)], [(0, "<", 50), (1, ">=", 0)], [(0, ">=", 10), (0, "<", 60), (1, "!=", 2), (2, "==", 1)],
    [(5, "==", 0)], [(2, ">=", -(2 ** 31) + 1)], [],
    # the ends of the int32 range and the missing code itself
    [(0, "<", -(2 ** 31))], [(0, ">", 2 ** 31 - 1)], [(1, "!=", -(2 ** 31))], [(1, "==", -(2 ** 31))], [(1, ">=", -(2 ** 31))],
    [(1, "<=", 2 ** 31 - 1)], [(1, "between", -(2 ** 31), 0)], [(1, "<=", -(2 ** 31))], [(1, "between", 3, 2 ** 31 - 1)],
]


@pytest.mark.parametrize("n", [1, 31, 32, 33, 127, 128, 129, 4097, 100_003])
def test_where_bitmap_is_bit_exact(n):
    from mlvectordb_b200 import DeviceShard
    rng = np.random.default_rng(n)
    s = DeviceShard(8, "l2")
    s.add(rng.standard_normal((n, 8)).astype(np.float32))
    cols = {0: rng.integers(0, 100, n).astype(np.int32), 1: rng.integers(-5, 5, n).astype(np.int32)}
    cols[1][rng.random(n) < 0.2] = MISSING
    half = n // 2
    s.set_column(0, cols[0])
    s.set_column(1, cols[1])
    col2 = np.full(n, MISSING, dtype=np.int32)          # column 2 written for the second half only
    col2[half:] = rng.integers(0, 3, n - half)
    if n - half:
        s.set_column(2, col2[half:], first_row=half)
    cols[2] = col2
    for c in (0, 1, 2):
        assert np.array_equal(s.get_column(c), cols[c])
    assert np.array_equal(s.get_column(9), np.full(n, MISSING, np.int32))      # never written
    for preds in PREDICATE_SETS:
        f = s.where(preds)
        want = exact.where_mask(cols, preds, n)
        assert np.array_equal(_bits(f.bitmap(), n), want), preds
        assert f.passing == int(want.sum())
        f.close()
    s.close()


def test_columns_follow_growth_deletes_and_compaction():
    from mlvectordb_b200 import DeviceShard
    rng = np.random.default_rng(11)
    dim = 16
    s = DeviceShard(dim, "l2", capacity=64)
    n1 = 50
    X1 = rng.standard_normal((n1, dim)).astype(np.float32)
    s.add(X1)
    c0 = rng.integers(0, 4, n1).astype(np.int32)
    s.set_column(0, c0)
    n2 = 5000                                            # forces the matrix (and lazily the column) to grow
    X2 = rng.standard_normal((n2, dim)).astype(np.float32)
    s.add(X2)
    col = np.concatenate([c0, np.full(n2, MISSING, np.int32)])
    assert np.array_equal(s.get_column(0), col)
    f = s.where([(0, "==", 1)])
    assert np.array_equal(_bits(f.bitmap(), n1 + n2), exact.where_mask({0: col}, [(0, "==", 1)], n1 + n2))
    f.close()
    c2 = rng.integers(0, 4, n2).astype(np.int32)
    s.set_column(0, c2, first_row=n1)
    col[n1:] = c2
    gone = rng.choice(n1 + n2, 1500, replace=False)
    s.mark_deleted(gone)
    mapping = s.compact()
    keep = mapping >= 0
    assert keep.sum() == n1 + n2 - 1500
    assert np.array_equal(s.get_column(0), col[keep])   # values moved with their rows
    X = np.concatenate([X1, X2])[keep]
    Q = rng.standard_normal((3, dim)).astype(np.float32)
    f = s.where([(0, "between", 1, 2)])
    mask = exact.where_mask({0: col[keep]}, [(0, "between", 1, 2)], int(keep.sum()))
    assert f.passing == int(mask.sum())
    assert _same(s.search(Q, 10, f), s.search(Q, 10, mask))
    L, D = exact.knn(X, Q, 10, "l2", allow=mask)
    d, r, c = s.search(Q, 10, f)
    for i in range(3):
        assert exact.check_topk_parity(r[i, :c[i]], d[i, :c[i]], L[i], D[i]) is None
    f.close()
    s.clear()
    s.add(X1)
    assert np.array_equal(s.get_column(0), np.full(n1, MISSING, np.int32))   # clear drops the columns
    s.close()


def test_where_errors():
    from mlvectordb_b200 import DeviceShard, _capi
    s = DeviceShard(4, "ip")
    s.add(np.ones((3, 4), np.float32))
    with pytest.raises(RuntimeError):
        s.set_column(_capi.MAX_COLUMNS, [1, 2, 3])
    with pytest.raises(RuntimeError):
        s.set_column(0, [1, 2, 3, 4])                   # beyond the stored rows
    with pytest.raises(RuntimeError):
        s.where([(0, "==", 1)] * (_capi.MAX_PREDICATES + 1))
    s.close()


@pytest.mark.parametrize("space", ["cosine", "l2"])
def test_index_filters_by_metadata_on_the_device(space):
    """``GpuIndex.add`` lifts the vectors' metadata into device columns; ``filter={...}`` searches equal the
    same search with the host-evaluated mask, bit for bit, and the oracle."""
    from mlvectordb_b200 import GpuIndex, GpuQueryProcessor, StoredVector, VectorDTO
    from mlvectordb_b200.columns import host_predicate
    from _refshim import InMemoryStorage
    n, dim, k = 3000, 32, 10
    X = synthetic.rows(21, 0, n, dim, scaled=True)
    rng = np.random.default_rng(2)
    mds = [{"bucket": int(rng.integers(0, 20)), "color": ["red", "green", "blue"][int(rng.integers(0, 3))]} if rng.random() < 0.9
           else {} for _ in range(n)]
    idx = GpuIndex(space=space)
    qp = GpuQueryProcessor(InMemoryStorage(), idx)
    vecs = [VectorDTO(values=X[i], metadata=mds[i]) for i in range(n)]
    qp.upsert_many(vecs[:1000], "ns")
    qp.upsert_many(vecs[1000:], "ns")
    assert sorted(idx.metadata_columns("ns")) == ["bucket", "color"]
    q = VectorDTO(values=synthetic.queries(21, 1, dim)[0], metadata={})
    ns = idx._ns["ns"]
    # ({"color": ("<", ...)}: an ORDERED constraint on a dictionary-coded key -- the column is re-coded by rank on the device)
    for cons in ({"color": "red"}, {"bucket": 3, "color": "blue"}, {"bucket": ("<", 5)}, {"bucket": ("between", 4, 9), "color": ("!=", "red")},
                 {"color": "purple"}, {"bucket": 3.5}, {"color": ("<", "h")}, {"color": ("between", "c", "red"), "bucket": (">=", 10)},
                 {"color": "red"}, {"color": (">", 7)}):
        assert idx.where("ns", cons) is not None
        mask = np.array([host_predicate(cons)(md) for md in mds])
        got = idx.search(q, k, "ns", space, filter=cons)
        ref = idx.search(q, k, "ns", space, filter=mask)
        assert [(r.vector_id, r.score) for r in got] == [(r.vector_id, r.score) for r in ref]
        assert len(got) == min(k, int(mask.sum()))
        L, D = exact.knn(X, np.asarray(q.values)[None, :], k, space, allow=mask)
        want_ids = [ns.uuid_of(int(l)) for l in L[0]]
        assert {r.vector_id for r in got} == set(want_ids) or exact.check_topk_parity(
            [ns.lookup()[r.vector_id.bytes] for r in got], [1 - r.score if space == "cosine" else r.score for r in got], L[0], D[0]) is None
        via_qp = qp.find_similar(q, k, "ns", space, filter=cons, enrich=False)
        assert [h["id"] for h in via_qp] == [r.vector_id for r in got]
    # constraints the device cannot decide fall back to the host inside the query processor
    assert idx.where("ns", {"color": None}) is None
    hits = qp.find_similar(q, k, "ns", space, filter={"color": None}, enrich=False)
    mask = np.array([md.get("color") is None for md in mds])
    ref = idx.search(q, k, "ns", space, filter=mask)
    assert [h["id"] for h in hits] == [r.vector_id for r in ref]
    # a mutation invalidates cached filters: a newly added matching row is found
    planted = VectorDTO(values=np.asarray(q.values), metadata={"color": "red", "bucket": 99})
    qp.insert(planted, "ns")
    top = qp.find_similar(q, 1, "ns", space, filter={"bucket": 99}, enrich=True)
    assert len(top) == 1 and top[0]["metadata"]["bucket"] == 99
    idx.close()


def test_bulk_columns_and_snapshot_round_trip(tmp_path):
    from mlvectordb_b200 import GpuIndex, VectorDTO
    n, dim, k = 20_000, 48, 10
    X = synthetic.rows(31, 0, n, dim, scaled=True)
    buckets = synthetic.buckets(31, 0, n).astype(np.int64)
    names = np.array(["a", "b", "c"])[np.arange(n) % 3]
    idx = GpuIndex(space="cosine", auto_compact=False)
    ids = idx.add_matrix(X, "big", columns={"bucket": buckets, "name": names})
    idx.add_matrix(X[:100, :16].copy(), "small")
    from uuid import UUID
    gone = [UUID(bytes=ids[i].tobytes()) for i in range(0, n, 7)]
    idx.remove(gone, "big")
    Q = synthetic.queries(31, 4, dim)
    cons = {"bucket": ("<", 10), "name": "b"}
    mask = (buckets < 10) & (names == "b")
    before = idx.search_batch(Q, k, "big", filter=cons)
    assert _same(before, idx.search_batch(Q, k, "big", filter=mask))
    before_plain = idx.search_batch(Q, k, "big")
    before_small = idx.search_batch(Q[:, :16].copy(), 5, "small")
    stored = idx._ns["big"].shard.export_rows()
    manifest = idx.save(str(tmp_path / "snap"))
    assert [m["name"] for m in manifest["namespaces"]] == ["big", "small"]
    idx.close()
    back = GpuIndex.load(str(tmp_path / "snap"))
    assert back._space == "cosine" and sorted(back.namespaces()) == ["big", "small"]
    assert np.array_equal(back._ns["big"].shard.export_rows(), stored)          # bit for bit, not re-normalised
    inf = back.info("big")
    assert inf["rows"] == n and inf["tombstones"] == len(gone)
    assert _same(back.search_batch(Q, k, "big"), before_plain)
    assert _same(back.search_batch(Q, k, "big", filter=cons), before)
    assert _same(back.search_batch(Q[:, :16].copy(), 5, "small"), before_small)
    res = back.search(VectorDTO(values=Q[0], metadata={}), k, "big", "cosine")
    assert all(r.vector_id not in set(gone) for r in res)
    # ids survived: removing by UUID still finds the row
    victim = res[0].vector_id
    back.remove([victim], "big")
    assert victim not in {r.vector_id for r in back.search(VectorDTO(values=Q[0], metadata={}), k, "big", "cosine")}
    back.close()


@pytest.mark.parametrize("dim", [70, 768])
def test_bulk_upload_through_pinned_staging_is_exact(dim):
    """``mlv_index_add`` of a large block goes through two pinned chunks filled by worker threads
    (``upload_rows_staged``); what lands in HBM equals the plain copy path and the input, bit for bit."""
    from mlvectordb_b200 import DeviceShard
    n = 300_000 if dim == 70 else 60_000          # 84 MB / 184 MB: several 48 MB chunks, ragged last chunk
    X = synthetic.rows(77, 0, n, dim)
    a = DeviceShard(dim, "l2")
    a.add(X[:1000])                               # small block: direct copy
    a.add(X[1000:])                               # staged
    b = DeviceShard(dim, "l2")
    b.set_tuning("staged_upload", 0)
    b.add(X)
    got = a.export_rows()
    assert np.array_equal(got, X) and np.array_equal(b.export_rows(), X)
    Q = synthetic.queries(78, 3, dim)
    assert _same(a.search(Q, 10), b.search(Q, 10))
    a.close()
    b.close()
