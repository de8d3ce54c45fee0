"""GPU parity: CUDA path (through the C ABI) vs the CPU oracle on identical seeded inputs.

Tolerance (BASELINE.json north_star): scores |a-b| <= 1e-5*max(|a|,|b|) + 1e-6,
id sets equal except swaps among candidates tied with the k-th score within that bound
(``oracle.exact.check_topk_parity``).  Integer outputs (rows, counts, tombstones, the synthetic
generator) are bit-exact.
"""
import numpy as np
import pytest

from oracle import cscan, exact, synthetic

pytestmark = pytest.mark.gpu


def _shard(dim, space, **kw):
    from mlvectordb_b200 import DeviceShard
    return DeviceShard(dim, space, **kw)


def _assert_knn(shard, X, Q, k, space, allow=None, filt=None):
    d, r, c = shard.search(Q, k, filt)
    L, D = exact.knn(X, Q, k, space, allow=allow)
    Xn = exact.normalize_rows(X) if space == "cosine" else X
    Qn = exact.normalize_rows(Q) if space == "cosine" else np.asarray(Q, np.float32).reshape(-1, X.shape[1])
    for i in range(len(L)):
        assert c[i] == len(L[i]), f"query {i}: count {c[i]} vs oracle {len(L[i])}"
        assert (r[i, c[i]:] == -1).all() and np.isinf(d[i, c[i]:]).all()
        msg = exact.check_topk_parity(
            r[i, :c[i]], d[i, :c[i]], L[i], D[i],
            all_ref_scores=lambda l, i=i: exact.distances(Xn[l:l + 1], Qn[i], space)[0])
        assert msg is None, f"{space} q{i}: {msg}"


@pytest.mark.parametrize("space", ["l2", "ip", "cosine"])
@pytest.mark.parametrize("n,dim", [(1, 4), (33, 3), (1000, 16), (5000, 37), (10000, 128), (4097, 768), (600, 1536)])
def test_knn_matches_oracle(space, n, dim):
    X = synthetic.rows(11 + dim, 0, n, dim, scaled=(space != "l2"))
    Q = synthetic.queries(11 + dim, 5, dim)
    Q[1] = X[n // 2]                      # planted exact match
    s = _shard(dim, space)
    assert s.add(X) == 0
    for k in (1, 10):
        _assert_knn(s, X, Q, min(k, n), space)
    s.close()


@pytest.mark.parametrize("k", [1, 7, 32, 33, 100, 257, 1000, 1024])
def test_k_sweep(k):
    n, dim = 3000, 24
    X = synthetic.rows(5, 0, n, dim)
    Q = synthetic.queries(5, 3, dim)
    s = _shard(dim, "l2")
    s.add(X)
    _assert_knn(s, X, Q, k, "l2")
    s.close()


@pytest.mark.parametrize("nq", [1, 2, 3, 4, 5, 8, 9, 17, 64])
@pytest.mark.parametrize("space", ["l2", "cosine"])
def test_query_batches(nq, space):
    n, dim = 2500, 96
    X = synthetic.rows(21, 0, n, dim, scaled=True)
    Q = synthetic.queries(21, nq, dim)
    s = _shard(dim, space)
    s.add(X)
    _assert_knn(s, X, Q, 10, space)
    s.close()


def test_k_larger_than_rows_pads():
    X = synthetic.rows(2, 0, 5, 8)
    s = _shard(8, "l2")
    s.add(X)
    d, r, c = s.search(synthetic.queries(2, 2, 8), 16)
    assert (c == 5).all()
    assert (r[:, 5:] == -1).all() and np.isinf(d[:, 5:]).all()
    assert sorted(r[0, :5].tolist()) == [0, 1, 2, 3, 4]
    s.close()


def test_duplicate_rows_tie_break_by_row():
    dim = 32
    X = synthetic.rows(9, 0, 500, dim)
    X[400] = X[17]
    X[250] = X[17]
    s = _shard(dim, "l2")
    s.add(X)
    d, r, c = s.search(X[17][None, :], 4)
    assert r[0, :3].tolist() == [17, 250, 400]
    assert (d[0, :3] == 0.0).all()
    s.close()


def test_empty_index_and_incremental_adds():
    dim = 20
    s = _shard(dim, "ip")
    d, r, c = s.search(synthetic.queries(1, 2, dim), 3)
    assert (c == 0).all() and (r == -1).all()
    X = synthetic.rows(1, 0, 2000, dim)
    # ragged appends crossing capacity doublings (1024 -> 2048)
    at = 0
    for step in (1, 7, 500, 516, 976):
        assert s.add(X[at:at + step]) == at
        at += step
    assert at == 2000 and s.rows == 2000
    _assert_knn(s, X, synthetic.queries(1, 3, dim), 10, "ip")
    s.close()


@pytest.mark.parametrize("space", ["l2", "cosine"])
def test_tombstones_and_compaction(space):
    n, dim = 3000, 48
    X = synthetic.rows(31, 0, n, dim, scaled=True)
    Q = synthetic.queries(31, 4, dim)
    Q[0] = X[100]
    s = _shard(dim, space)
    s.add(X)
    rng = np.random.default_rng(0)
    dead = rng.choice(n, size=700, replace=False)
    dead[0] = 100                                   # the planted match dies
    dead = np.unique(dead)
    assert s.mark_deleted(dead) == len(dead)
    assert s.mark_deleted(dead[:10]) == 0           # idempotent
    assert s.mark_deleted(np.array([n + 5])) == 0   # out of range ignored
    assert s.live == n - len(dead)
    allow = np.ones(n, bool)
    allow[dead] = False
    _assert_knn(s, X, Q, 10, space, allow=allow)
    # compaction renumbers survivors in order
    mapping = s.compact()
    keep = np.flatnonzero(allow)
    assert (mapping[dead] == -1).all()
    assert (mapping[keep] == np.arange(len(keep))).all()
    assert s.rows == len(keep) and s.live == len(keep)
    _assert_knn(s, X[keep], Q, 10, space)
    # and the matrix keeps growing after compaction
    X2 = synthetic.rows(32, 0, 100, dim, scaled=True)
    assert s.add(X2) == len(keep)
    _assert_knn(s, np.concatenate([X[keep], X2]), Q, 10, space)
    s.close()


@pytest.mark.parametrize("sel", [0.0, 0.01, 0.1, 0.5, 1.0])
def test_filter_bitmap(sel):
    n, dim = 6000, 64
    X = synthetic.rows(41, 0, n, dim)
    Q = synthetic.queries(41, 3, dim)
    mask = synthetic.buckets(41, 0, n) < int(100 * sel)
    s = _shard(dim, "l2")
    s.add(X)
    _assert_knn(s, X, Q, 10, "l2", allow=mask, filt=mask)
    # packed words are accepted too, and combine with tombstones
    s.mark_deleted(np.arange(0, n, 3))
    both = mask.copy()
    both[::3] = False
    _assert_knn(s, X, Q, 10, "l2", allow=both, filt=synthetic.bitmap_from_mask(mask))
    s.close()


@pytest.mark.parametrize("space", ["l2", "ip", "cosine"])
def test_range_search(space):
    n, dim = 5000, 40
    X = synthetic.rows(51, 0, n, dim, scaled=True)
    Q = synthetic.queries(51, 3, dim)
    Q[2] = X[77]
    s = _shard(dim, space)
    s.add(X)
    # radius = oracle's 50th smallest distance of query 0
    _, D = exact.knn(X, Q[:1], 50, space)
    radius = float(D[0][-1])
    res = s.range_search(Q, radius, max_hits=8)     # forces the retry path
    L, Dr = exact.range_search(X, Q, radius, space)
    tol = 1e-5 * abs(radius) + 1e-6
    Xn = exact.normalize_rows(X) if space == "cosine" else X
    Qn = exact.normalize_rows(Q) if space == "cosine" else Q
    for i, (d, r) in enumerate(res):
        got, ref = set(r.tolist()), set(L[i].tolist())
        # rows may differ only where the oracle distance sits on the radius within tolerance
        for l in got ^ ref:
            assert abs(float(exact.distances(Xn[l:l + 1], Qn[i], space)[0]) - radius) <= 2 * tol
        assert (np.diff(d) >= 0).all()
        common = [j for j, l in enumerate(r.tolist()) if l in ref]
        ref_pos = {l: j for j, l in enumerate(L[i].tolist())}
        a = d[common]
        b = np.array([Dr[i][ref_pos[r[j]]] for j in common])
        assert exact.scores_close(a, b).all()
    s.close()


def test_synthetic_generator_bit_exact():
    for dim, scaled in ((5, False), (128, True), (770, True)):
        s = _shard(dim, "l2")
        s.add_synthetic(42, 1000, 300, scaled)
        got = s.get_rows(np.arange(300))
        assert np.array_equal(got, synthetic.rows(42, 1000, 300, dim, scaled))
        assert np.array_equal(got, cscan.fill_synthetic(42, 1000, 300, dim, scaled))
        s.close()


def test_cosine_rows_stored_normalised():
    dim = 100
    X = synthetic.rows(61, 0, 64, dim, scaled=True)
    s = _shard(dim, "cosine")
    s.add(X)
    got = s.get_rows(np.arange(64))
    np.testing.assert_allclose(got, exact.normalize_rows(X), rtol=2e-6, atol=1e-8)
    s.close()


def test_row_base_offsets_results():
    X = synthetic.rows(3, 0, 100, 8)
    s = _shard(8, "l2", row_base=5_000_000_000)
    s.add(X)
    d, r, c = s.search(X[10][None], 1)
    assert r[0, 0] == 5_000_000_010
    s.close()


def test_dimension_mismatch_and_bad_k_raise():
    s = _shard(8, "l2")
    s.add(synthetic.rows(3, 0, 10, 8))
    with pytest.raises(ValueError):
        s.search(np.zeros((1, 7), np.float32), 1)
    with pytest.raises(RuntimeError):
        s.search(np.zeros((1, 8), np.float32), 2000)   # > MLV_MAX_K
    s.close()


# ---- full-size, size-independent properties (oracle too slow to hold 1M x 768 per test) -----
def test_large_planted_matches_and_shard_merge_property():
    n, dim, k = 1_000_000, 768, 10
    s = _shard(dim, "cosine", capacity=n)
    s.add_synthetic(42, 0, n, True)
    planted = np.array([0, 123_456, 999_999])
    Q = synthetic.rows(42, 0, 1, dim, True)
    Q = np.concatenate([synthetic.rows(42, int(p), 1, dim, True) for p in planted] + [synthetic.queries(42, 2, dim)])
    d, r, c = s.search(Q, k)
    assert (c == k).all()
    assert r[:3, 0].tolist() == planted.tolist()           # a stored row is its own nearest neighbour
    np.testing.assert_allclose(d[:3, 0], 0.0, atol=2e-6)   # cosine distance of a row to itself
    assert (np.diff(d, axis=1) >= 0).all()
    # exactness at full size: oracle on the same generator, streamed in chunks
    def chunks():
        for s0 in range(0, n, 100_000):
            yield s0, cscan.fill_synthetic(42, s0, 100_000, dim, True)
    L, D = exact.knn_stream(chunks(), Q[3:], k, "cosine")
    for i in range(2):
        msg = exact.check_topk_parity(r[3 + i], d[3 + i], L[i], D[i])
        assert msg is None, msg
    # top-k of a union of row shards == merge of the shards' top-k (the multi-GPU identity)
    halves = []
    for base in (0, n // 2):
        h = _shard(dim, "cosine", capacity=n // 2, row_base=base)
        h.add_synthetic(42, base, n // 2, True)
        halves.append(h.search(Q, k))
        h.close()
    md = np.concatenate([halves[0][0], halves[1][0]], axis=1)
    mr = np.concatenate([halves[0][1], halves[1][1]], axis=1)
    for i in range(Q.shape[0]):
        order = np.lexsort((mr[i], md[i]))[:k]
        assert mr[i][order].tolist() == r[i].tolist()
        assert np.array_equal(md[i][order], d[i])
    s.close()


def test_range_search_hit_lists_longer_than_one_sort_block():
    """> 8192 hits: the device leaves them unordered, the host entry point orders them -- same contract."""
    n, dim = 40_000, 16
    X = synthetic.rows(77, 0, n, dim)
    Q = synthetic.queries(77, 2, dim)
    s = _shard(dim, "l2")
    s.add(X)
    ds = np.sort(exact.distances(X, Q[0], "l2"))
    radius = float((ds[12_000] + ds[12_001]) / 2)
    (d0, r0), (d1, r1) = s.range_search(Q, radius, max_hits=64)   # first call reports totals, shim retries
    assert len(r0) == 12_001
    L, D = exact.range_search(X, Q, radius, "l2")
    for got_r, got_d, ref_r, ref_d in ((r0, d0, L[0], D[0]), (r1, d1, L[1], D[1])):
        # ids may differ only for rows whose oracle distance is within tolerance of the radius (each adjudicated)
        tol = 1e-5 * abs(radius) + 1e-6
        qi = 0 if got_r is r0 else 1
        odd = sorted(set(got_r.tolist()) ^ set(np.asarray(ref_r).tolist()))
        assert abs(len(got_r) - len(ref_r)) <= len(odd)
        for row in odd:
            assert abs(float(exact.distances(X[row:row + 1], Q[qi], "l2")[0]) - radius) <= 2 * tol, f"row {row} is not at the radius"
        assert np.all(np.diff(got_d) >= 0), "hits are not ascending"
        keep = ~np.isin(got_r, odd)
        ref_keep = ~np.isin(np.asarray(ref_r), odd)
        assert got_r[keep].tolist() == np.asarray(ref_r)[ref_keep].tolist() or exact.scores_close(got_d[keep], np.asarray(ref_d)[ref_keep]).all()
    s.close()


def test_searches_on_several_streams_share_one_handle():
    """Device entry point on up to four streams at once (per-stream scratch lanes), then more streams
    than lanes (a lane is recycled after its previous stream drained): every result equals the
    single-stream result."""
    import torch

    n, dim, k = 200_000, 64, 10
    s = _shard(dim, "cosine", capacity=n)
    s.add_synthetic(5, 0, n, True)
    Q = synthetic.queries(6, 24, dim)
    ref_d, ref_r, ref_c = s.search(Q, k)
    dev = torch.device("cuda", 0)
    Qd = torch.from_numpy(Q).to(dev)
    for n_streams in (2, 4, 6):
        streams = [torch.cuda.Stream(dev) for _ in range(n_streams)]
        outs = []
        torch.cuda.synchronize()
        for j in range(Q.shape[0]):
            st = streams[j % n_streams]
            d = torch.empty((1, k), dtype=torch.float32, device=dev)
            r = torch.empty((1, k), dtype=torch.int64, device=dev)
            c = torch.empty((1,), dtype=torch.int32, device=dev)
            with torch.cuda.stream(st):
                s.search_device(Qd[j:j + 1].data_ptr(), 1, k, d.data_ptr(), r.data_ptr(), c.data_ptr(), stream=st.cuda_stream)
            outs.append((d, r, c))
        torch.cuda.synchronize()
        for j, (d, r, c) in enumerate(outs):
            assert int(c.item()) == ref_c[j]
            assert np.array_equal(r.cpu().numpy()[0], ref_r[j]), f"{n_streams} streams, query {j}"
            assert np.array_equal(d.cpu().numpy()[0], ref_d[j])
    s.close()


@pytest.mark.parametrize("space", ["l2", "ip", "cosine"])
@pytest.mark.parametrize("n,dim,k", [(10_000, 128, 10), (3000, 37, 5), (50_000, 768, 27), (700, 2048, 3), (200, 5, 1)])
def test_single_query_latency_path_equals_the_staged_path(space, n, dim, k):
    """``mlv_index_search`` with one query takes the one-launch path (raw query in the launch parameters, normalised in
    the kernel, results written to mapped pinned memory, flag polled): bit-identical to the staged path
    (H2D + prep_queries_kernel + scan + D2H), to a batch holding the same query, and within tolerance of the oracle."""
    X = synthetic.rows(70 + dim, 0, n, dim, scaled=True)
    Q = synthetic.queries(70 + dim, 6, dim) * np.float32(3.0)     # un-normalised queries: the in-kernel normalisation matters
    Q[2] = X[n // 3]
    s = _shard(dim, space)
    s.add(X)
    fast = [s.search(Q[i:i + 1], k) for i in range(6)]
    s.set_tuning("fast_host", 0)
    staged = [s.search(Q[i:i + 1], k) for i in range(6)]
    s.set_tuning("fast_host", 1)
    batch = s.search(Q[:4], k)
    for i in range(6):
        assert all(np.array_equal(a, b) for a, b in zip(fast[i], staged[i])), f"query {i}"
    for i in range(4):
        assert np.array_equal(fast[i][1][0], batch[1][i]) and np.array_equal(fast[i][0][0], batch[0][i])
    L, D = exact.knn(X, Q, k, space)
    for i in range(6):
        msg = exact.check_topk_parity(fast[i][1][0], fast[i][0][0], L[i], D[i])
        assert msg is None, msg
    # tombstones and a bound prepared filter ride the same path
    s.mark_deleted(np.arange(0, n, 5))
    mask = np.arange(n) % 3 != 0
    f = s.prepare_filter(mask)
    a = s.search(Q[1:2], k, f)
    s.set_tuning("fast_host", 0)
    b = s.search(Q[1:2], k, f)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    allow = mask & (np.arange(n) % 5 != 0)
    Lf, Df = exact.knn(X, Q[1:2], k, space, allow=allow)
    assert exact.check_topk_parity(a[1][0, :a[2][0]], a[0][0, :a[2][0]], Lf[0], Df[0]) is None
    f.close()
    s.close()


def test_single_query_latency_path_many_calls_and_k_beyond_the_fused_tail():
    """Flag values advance call by call; k too large for the fused tail falls back to the staged path silently."""
    n, dim = 20_000, 64
    X = synthetic.rows(5, 0, n, dim)
    s = _shard(dim, "l2")
    s.add(X)
    Q = synthetic.queries(6, 50, dim)
    ref = s.search(Q, 7)
    for i in range(50):
        d, r, c = s.search(Q[i:i + 1], 7)
        assert np.array_equal(r[0], ref[1][i]) and np.array_equal(d[0], ref[0][i])
    big = s.search(Q[:1], 200)                # 148 * 200 keys: separate select kernel
    L, D = exact.knn(X, Q[:1], 200, "l2")
    assert exact.check_topk_parity(big[1][0], big[0][0], L[0], D[0]) is None
    s.close()
