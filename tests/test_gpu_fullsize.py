"""GPU parity at BASELINE.json's FULL sizes through size-independent properties (the CPU oracle cannot hold
30 GB per test): paths tied to each other bit for bit, planted matches, sortedness, range/kNN consistency.

  configs[2]  10M x 768 l2 k=100, 4096-query batch   tensor-core tiers == scan on sampled queries, planted rows first
  configs[3]  10M x 384 cosine k=10, 1 % / 10 % filter device predicate == host mask; gather == stream+mask
  configs[4]  100M x 128 ip, one GPU's shard of eight (12.5M rows): kNN and range search agree; a row is its own match
"""
import numpy as np
import pytest

from oracle import cscan, exact, synthetic

pytestmark = pytest.mark.gpu


def _same(a, b):
    return all(np.array_equal(x, y, equal_nan=True) for x, y in zip(a, b))


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
    # ... and the reference-form arithmetic itself on the rows that were returned (CPU oracle, query 7)
    rows7 = r[7]
    X7 = np.concatenate([cscan.fill_synthetic(42, int(x), 1, dim, False) for x in rows7[:10]])
    want = exact.distances(X7, Q[7], "l2")
    assert np.allclose(d[7, :10], want, rtol=1e-5, atol=1e-6)
    s.close()


@pytest.mark.parametrize("cut", [1, 10])
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
    # scores are the reference's arithmetic (1 - dot, fp32) on the returned rows
    X0 = np.concatenate([cscan.fill_synthetic(42, int(x), 1, dim, True) for x in r[0, :10]])
    assert np.allclose(d[0, :10], exact.distances(X0, Q[0], "ip"), rtol=1e-5, atol=1e-6)
    s.close()
