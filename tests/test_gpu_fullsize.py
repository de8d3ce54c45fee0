"""GPU parity at BASELINE.json's FULL sizes.  Two layers:

* the STREAMED ORACLE (``oracle.cscan.knn_synthetic`` / ``range_synthetic``: the C restatement regenerating the
  synthetic rows on the fly, so 10M x 768 costs seconds of CPU and no memory) with the full parity rule
  ``oracle.exact.check_topk_parity`` -- this is what catches an error common to every CUDA path (row-index
  overflow, a lost tail tile, a scheduler hole) at the sizes BASELINE quotes (SURVEY.md 8d "run at full N");
* size-independent properties on top: paths tied to each other bit for bit, planted matches, sortedness,
  range/kNN consistency.

  headline    10M x 768 cosine k=10, batch-1         oracle on 6 queries (random, planted, last row, duplicated-row tie)
  configs[2]  10M x 768 l2 k=100, 4096-query batch   oracle on 8 queries of the batch; tensor-core tiers == scan; planted rows first
  configs[3]  10M x 384 cosine k=10, 1 / 10 / 50 %   oracle restricted to the passing rows; device predicate == host mask; gather == stream+mask
  configs[4]  100M x 128 ip, one GPU's shard of eight (12.5M rows): oracle kNN (k = 10, 100) + range lists; kNN and range agree
  side run    1M x 768 cosine with 10 % tombstones   oracle restricted to the live rows
"""
import numpy as np
import pytest

from oracle import cscan, exact, synthetic

pytestmark = pytest.mark.gpu


def _same(a, b):
    return all(np.array_equal(x, y, equal_nan=True) for x, y in zip(a, b))


@pytest.fixture(autouse=True, scope="module")
def _all_cores():
    cscan.use_all_cores()


def test_headline_10Mx768_cosine_k10_against_the_streamed_oracle():
    """BASELINE.json metric config: batch-1 queries, every one a full pass.  Queries: random, a planted stored row,
    the LAST row (tail tile), and a row whose duplicate is appended after the synthetic block (bit-exact tie)."""
    from mlvectordb_b200 import DeviceShard
    n, dim, k = 10_000_000, 768, 10
    s = DeviceShard(dim, "cosine", capacity=n + 8)
    s.add_synthetic(42, 0, n, True)
    dup_of = 7_654_321
    dup = cscan.fill_synthetic(42, dup_of, 1, dim, True)
    assert s.add(dup) == n                                   # row n is a copy of row dup_of (normalised on the device)
    Q = np.concatenate([synthetic.queries(43, 3, dim), cscan.fill_synthetic(42, 4_999_999, 1, dim, True),
                        cscan.fill_synthetic(42, n - 1, 1, dim, True), dup])
    got = [s.search(Q[i:i + 1], k) for i in range(Q.shape[0])]          # batch-1, like the bench
    d = np.concatenate([g[0] for g in got]); r = np.concatenate([g[1] for g in got]); c = np.concatenate([g[2] for g in got])
    # oracle over the synthetic block; the appended duplicate scores exactly like its original
    L, D, C = cscan.knn_synthetic(42, 0, n, dim, True, Q, k, "cosine")
    for i in range(Q.shape[0]):
        pos = np.flatnonzero(L[i] == dup_of)
        if pos.size:       # insert (same distance, label n) right after the original, drop the last
            j = int(pos[0]) + 1
            L[i] = np.concatenate([L[i, :j], [n], L[i, j:]])[:k]
            D[i] = np.concatenate([D[i, :j], [D[i, j - 1]], D[i, j:]])[:k]
        msg = exact.check_topk_parity(r[i, :c[i]], d[i, :c[i]], L[i], D[i],
                                      all_ref_scores=lambda l, i=i: float(cscan.distances_synthetic(42, [dup_of if l == n else l], dim, True, Q[i], "cosine")[0]))
        assert c[i] == k and msg is None, f"query {i}: {msg}"
    assert r[3, 0] == 4_999_999 and r[4, 0] == n - 1
    assert r[5, :2].tolist() == [dup_of, n] and d[5, 0] == d[5, 1]      # duplicated rows tie bit for bit, ordered by row
    # the same queries as one small batch (multi-query scan pass) return the same bits
    assert _same(s.search(Q[:4], k), (d[:4], r[:4], c[:4]))
    s.close()


def test_tombstones_10pct_at_1Mx768_against_the_streamed_oracle():
    """SURVEY.md 8d side run: 10 % random deletes; tombstoned rows never come back, the rest matches the oracle."""
    from mlvectordb_b200 import DeviceShard
    n, dim, k = 1_000_000, 768, 10
    s = DeviceShard(dim, "cosine", capacity=n)
    s.add_synthetic(42, 0, n, True)
    gone = np.random.default_rng(5).choice(n, n // 10, replace=False)
    assert s.mark_deleted(gone) == n // 10
    live = np.ones(n, bool)
    live[gone] = False
    Q = np.concatenate([synthetic.queries(47, 3, dim), cscan.fill_synthetic(42, int(gone[0]), 1, dim, True)])
    d, r, c = s.search(Q, k)
    assert live[r].all()
    assert cscan.check_knn_synthetic(r, d, c, 42, 0, n, dim, True, Q, k, "cosine", allow=live) is None
    s.close()


def test_config3_batch_path_at_full_size():
    from mlvectordb_b200 import DeviceShard
    n, dim, k, nq = 10_000_000, 768, 100, 4096
    s = DeviceShard(dim, "l2", capacity=n)
    s.add_synthetic(42, 0, n, False)
    Q = synthetic.queries(43, nq, dim)
    planted = [0, 4_999_999, 9_999_999]
    for i, row in enumerate(planted):
        Q[i] = cscan.fill_synthetic(42, row, 1, dim, False)[0]
    before = s.gemm_stats()
    d, r, c = s.search(Q, k)
    after = s.gemm_stats()
    assert after["searches"] == before["searches"] + 1 and after["queries"] - before["queries"] == nq
    assert after["fallback_queries"] == before["fallback_queries"], "benign data must not need the scan"
    assert (c == k).all() and (np.diff(d, axis=1) >= 0).all()
    assert r[:3, 0].tolist() == planted and (d[:3, 0] == 0).all()       # squared L2 of a row to itself
    assert all(len(set(row.tolist())) == k for row in r[::257])
    # the exact scan on a sample of the batch: bit-identical rows, scores, counts
    pick = np.array([0, 1, 2, 7, 1000, 2048, 4095])
    s.set_tuning("gemm", 0)
    ref = s.search(Q[pick], k)
    assert _same((d[pick], r[pick], c[pick]), ref)
    # ... and the streamed oracle at full N on 8 queries of the batch (planted, first / last tile of the batch, middle)
    pick8 = np.array([0, 2, 7, 255, 256, 1000, 2049, 4095])
    assert cscan.check_knn_synthetic(r[pick8], d[pick8], c[pick8], 42, 0, n, dim, False, Q[pick8], k, "l2") is None
    # ... and the reference-form arithmetic itself on the rows that were returned (CPU oracle, query 7)
    rows7 = r[7]
    X7 = np.concatenate([cscan.fill_synthetic(42, int(x), 1, dim, False) for x in rows7[:10]])
    want = exact.distances(X7, Q[7], "l2")
    assert np.allclose(d[7, :10], want, rtol=1e-5, atol=1e-6)
    s.close()


@pytest.mark.parametrize("cut", [1, 10, 50])
def test_config4_filters_at_full_size(cut):
    from mlvectordb_b200 import DeviceShard
    n, dim, k = 10_000_000, 384, 10
    s = DeviceShard(dim, "cosine", capacity=n)
    s.add_synthetic(42, 0, n, True)
    buckets = synthetic.buckets(44, 0, n)
    s.set_column(0, buckets)
    mask = buckets < cut
    f = s.where([(0, "<", cut)])
    assert f.passing == int(mask.sum())
    assert np.array_equal(f.bitmap(), np.packbits(mask, bitorder="little").view(np.uint32))   # n % 32 == 0
    Q = synthetic.queries(45, 4, dim)
    Q[0] = cscan.fill_synthetic(42, int(np.flatnonzero(mask)[1234]), 1, dim, True)[0]            # a passing stored row
    s.set_tuning("gather", 1)
    gathered = s.search(Q, k, f)
    s.set_tuning("gather", 0)
    masked = s.search(Q, k, f)
    s.set_tuning("gather", -1)
    assert _same(gathered, masked) and _same(gathered, s.search(Q, k, mask))
    d, r, c = gathered
    assert (c == k).all() and mask[r].all() and (np.diff(d, axis=1) >= 0).all()
    assert r[0, 0] == np.flatnonzero(mask)[1234] and abs(d[0, 0]) < 2e-6
    # the streamed oracle restricted to the passing rows, at full N
    assert cscan.check_knn_synthetic(r, d, c, 42, 0, n, dim, True, Q, k, "cosine", allow=mask) is None
    # against the unfiltered search: every unfiltered hit that passes the filter is in the filtered list's prefix
    ud, ur, uc = s.search(Q, 100)
    for i in range(4):
        keep = ur[i][mask[ur[i]]][:k]
        assert r[i, : len(keep)].tolist() == keep.tolist()
    f.close()
    s.close()


def test_config5_shard_knn_and_range_agree():
    from mlvectordb_b200 import DeviceShard
    n, dim, k = 12_500_000, 128, 10      # one of eight shards of 100M x 128
    base = 3 * n                          # the fourth shard: generator rows and row_base 37.5M ..
    s = DeviceShard(dim, "ip", capacity=n, row_base=base)
    s.add_synthetic(42, base, n, True)
    Q = synthetic.queries(46, 3, dim)
    d, r, c = s.search(Q, 100)
    assert (c == 100).all() and (r >= base).all() and (r < base + n).all() and (np.diff(d, axis=1) >= 0).all()
    for i in range(3):
        radius = float(d[i, 99])
        (hd, hr), = s.range_search(Q[i:i + 1], radius)
        ties = int((d[i] == d[i, 99]).sum())          # rows tied with the 100th may extend the range result
        assert len(hr) >= 100 and len(hr) < 100 + 64
        assert hr[: 100 - ties + 1].tolist() == r[i, : 100 - ties + 1].tolist()
        assert np.array_equal(hd[:100], d[i]) and (hd <= radius).all()
    # streamed oracle over this shard's 12.5M generator rows: kNN (k = 10 and 100) and the range lists
    assert cscan.check_knn_synthetic(r, d, c, 42, base, n, dim, True, Q, 100, "ip") is None
    d10, r10, c10 = s.search(Q, 10)
    assert cscan.check_knn_synthetic(r10, d10, c10, 42, base, n, dim, True, Q, 10, "ip") is None
    for i in range(3):
        radius = float(d[i, 99])
        (hd, hr), = s.range_search(Q[i:i + 1], radius)
        (ol, od), = cscan.range_synthetic(42, base, n, dim, True, Q[i:i + 1], radius, "ip")
        # ids may differ only where the oracle's distance is within tolerance of the radius
        tol = 1e-5 * abs(radius) + 1e-6
        for odd in set(hr.tolist()) ^ set(ol.tolist()):
            assert abs(float(cscan.distances_synthetic(42, [odd], dim, True, Q[i], "ip")[0]) - radius) <= 2 * tol
        common = min(len(hr), len(ol))
        assert abs(len(hr) - len(ol)) <= 2 and exact.scores_close(hd[:common - 2], od[:common - 2]).all()
    # scores are the reference's arithmetic (1 - dot, fp32) on the returned rows
    X0 = np.concatenate([cscan.fill_synthetic(42, int(x), 1, dim, True) for x in r[0, :10]])
    assert np.allclose(d[0, :10], exact.distances(X0, Q[0], "ip"), rtol=1e-5, atol=1e-6)
    s.close()
