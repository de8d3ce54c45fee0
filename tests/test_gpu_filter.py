"""GPU parity of filtered search (BASELINE.json configs[3]): the gather path (only passing rows are
copied out of HBM) and the stream+mask path return bit-identical results, and both match the
oracle's hnswlib-0.8 ``filter=`` semantics (``oracle.exact.knn(allow=...)``)."""
import numpy as np
import pytest

from oracle import exact, synthetic

pytestmark = pytest.mark.gpu


def _shard(dim, space, **kw):
    from mlvectordb_b200 import DeviceShard
    return DeviceShard(dim, space, **kw)


def _oracle_check(got, X, Q, k, space, allow):
    d, r, c = got
    L, D = exact.knn(X, Q, k, space, allow=allow)
    Xn = exact.normalize_rows(X) if space == "cosine" else X
    Qn = exact.normalize_rows(Q) if space == "cosine" else np.asarray(Q, np.float32)
    for i in range(len(L)):
        assert c[i] == len(L[i])
        msg = exact.check_topk_parity(r[i, :c[i]], d[i, :c[i]], L[i], D[i],
                                      all_ref_scores=lambda l, i=i: exact.distances(Xn[l:l + 1], Qn[i], space)[0])
        assert msg is None, msg


def _same(a, b):
    return all(np.array_equal(x, y, equal_nan=True) for x, y in zip(a, b))


@pytest.mark.parametrize("space", ["l2", "cosine"])
@pytest.mark.parametrize("sel", [0.0005, 0.01, 0.1, 0.5, 1.0])
def test_gather_equals_mask_equals_oracle(space, sel):
    n, dim, k = 60_000, 96, 10
    X = synthetic.rows(3, 0, n, dim, scaled=True)
    Q = synthetic.queries(3, 5, dim)
    rng = np.random.default_rng(5)
    mask = rng.random(n) < sel
    s = _shard(dim, space)
    s.add(X)
    s.set_tuning("gather", 0)
    masked = s.search(Q, k, mask)
    s.set_tuning("gather", 1)
    gathered = s.search(Q, k, mask)
    assert _same(masked, gathered)
    _oracle_check(gathered, X, Q, k, space, mask)
    pf = s.prepare_filter(mask)
    assert pf.passing == int(mask.sum())
    s.set_tuning("gather", -1)
    assert _same(s.search(Q, k, pf), gathered)       # auto: gather when selective, mask when dense
    s.set_tuning("gather", 1)
    assert _same(s.search(Q, k, pf), gathered)
    pf.close()
    s.close()


def test_prepared_filter_follows_adds_and_deletes_and_dies_with_compaction():
    n, dim, k = 20_000, 48, 10
    X = synthetic.rows(7, 0, n + 3000, dim)
    Q = synthetic.queries(7, 4, dim)
    rng = np.random.default_rng(1)
    mask = rng.random(n) < 0.05
    s = _shard(dim, "l2")
    s.add(X[:n])
    pf = s.prepare_filter(mask)
    dead = np.flatnonzero(mask)[::3]
    s.mark_deleted(dead.astype(np.uint64))
    allow = mask.copy()
    allow[dead] = False
    assert pf.passing == int(allow.sum())
    _oracle_check(s.search(Q, k, pf), X[:n], Q, k, "l2", allow)
    s.add(X[n:])                                     # rows appended later do not pass
    allow2 = np.zeros(n + 3000, bool)
    allow2[:n] = allow
    _oracle_check(s.search(Q, k, pf), X, Q, k, "l2", allow2)
    s.compact()
    with pytest.raises(RuntimeError):
        s.search(Q, k, pf)
    pf.close()
    s.close()


def test_range_search_with_gather_filter():
    n, dim = 30_000, 32
    X = synthetic.rows(9, 0, n, dim)
    Q = synthetic.queries(9, 3, dim)
    rng = np.random.default_rng(2)
    mask = rng.random(n) < 0.2
    s = _shard(dim, "l2")
    s.add(X)
    ds = np.sort(exact.distances(X[mask], Q[0], "l2"))
    radius = float((ds[40] + ds[41]) / 2)           # between two hits: summation order cannot move the count
    s.set_tuning("gather", 0)
    a = s.range_search(Q, radius, mask)
    s.set_tuning("gather", 1)
    b = s.range_search(Q, radius, mask)
    for (da, ra), (db, rb) in zip(a, b):
        assert np.array_equal(ra, rb) and np.array_equal(da, db)
    assert len(b[0][1]) == 41 and mask[b[0][1]].all()
    s.close()


def test_batched_queries_with_filter_take_both_paths():
    """nq = 300: tensor-core path masks in its epilogue (bitmap), scan path gathers -- identical."""
    n, dim, k = 40_000, 64, 10
    X = synthetic.rows(11, 0, n, dim, scaled=True)
    Q = synthetic.queries(11, 300, dim)
    mask = np.random.default_rng(3).random(n) < 0.1
    s = _shard(dim, "cosine")
    s.add(X)
    s.set_tuning("gemm", 0)
    a = s.search(Q, k, mask)
    pf = s.prepare_filter(mask)
    s.set_tuning("gemm", 1)
    b = s.search(Q, k, mask)
    c = s.search(Q, k, pf)
    assert _same(a, b) and _same(a, c)
    _oracle_check((a[0][:12], a[1][:12], a[2][:12]), X, Q[:12], k, "cosine", mask)
    pf.close()
    s.close()
