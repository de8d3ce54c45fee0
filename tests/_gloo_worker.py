"""One rank of the world_size-2 gloo test (launched by tests/test_sharded_gloo.py as a subprocess).

usage: _gloo_worker.py RANK WORLD PORT TOTAL_ROWS DIM K SPACE
The two device steps of ``ShardedIndex`` are injected: the local scan is played by the oracle on
this rank's row block, the merge kernel by a numpy restatement of ``merge_pairs_kernel``'s
contract; the collective plumbing (partitioning, all-gather layout) is the product's.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def numpy_merge(gd, gr, k):
    """[G, nq, k] candidates -> global top-k by (distance, list position); -1 rows are empty slots."""
    G, nq, _ = gd.shape
    d = gd.permute(1, 0, 2).reshape(nq, G * k).numpy()
    r = gr.permute(1, 0, 2).reshape(nq, G * k).numpy()
    out_d = np.full((nq, k), np.inf, np.float32)
    out_r = np.full((nq, k), -1, np.int64)
    out_c = np.zeros(nq, np.int32)
    for i in range(nq):
        valid = np.flatnonzero(r[i] >= 0)
        order = valid[np.lexsort((valid, d[i][valid]))][:k]
        out_d[i, :len(order)], out_r[i, :len(order)], out_c[i] = d[i][order], r[i][order], len(order)
    return torch.from_numpy(out_d), torch.from_numpy(out_r), torch.from_numpy(out_c)


def main():
    rank, world, port, total_rows, dim, k = (int(x) for x in sys.argv[1:7])
    space = sys.argv[7]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mlvectordb_b200.sharded import ShardedIndex, shard_range
    from oracle import exact, synthetic

    lo, hi = shard_range(total_rows, rank, world)
    X = synthetic.rows(5, lo, hi - lo, dim, scaled=True)

    def local_search(q, kk):
        nq = q.shape[0]
        d = np.full((nq, kk), np.inf, np.float32)
        r = np.full((nq, kk), -1, np.int64)
        c = np.zeros(nq, np.int32)
        if hi > lo:
            L, D = exact.knn(X, q.numpy(), kk, space)
            for i in range(nq):
                m = len(L[i])
                d[i, :m], r[i, :m], c[i] = D[i], L[i] + lo, m
        return torch.from_numpy(d), torch.from_numpy(r), torch.from_numpy(c)

    def local_range(q, radius):
        if hi == lo:
            return [(np.empty(0, np.float32), np.empty(0, np.int64)) for _ in range(q.shape[0])]
        L, D = exact.range_search(X, q, radius, space)
        return [(np.asarray(D[i], np.float32), np.asarray(L[i], np.int64) + lo) for i in range(q.shape[0])]

    def order_hits(d, r):   # stands in for mlv_index_order_pairs_device: (distance, row) ascending, padding (row -1) last
        dn, rn = d.numpy(), r.numpy()
        order = np.lexsort((rn, dn, rn < 0))
        return torch.from_numpy(dn[order]), torch.from_numpy(rn[order])

    idx = ShardedIndex(dim, space, total_rows, device=None, local_search=local_search, merge=numpy_merge,
                       local_range=local_range, order_hits=order_hits)
    assert (idx.lo, idx.hi) == (lo, hi) and idx.world == world
    Q = synthetic.queries(5, 4, dim)
    d, r, c = idx.search_device(torch.from_numpy(Q), k)
    gathered = [torch.empty_like(r) for _ in range(world)]
    dist.all_gather(gathered, r)
    assert all(torch.equal(g, r) for g in gathered), "ranks disagree on the global answer"
    full = synthetic.rows(5, 0, total_rows, dim, scaled=True)
    L, D = exact.knn(full, Q, k, space)
    for i in range(4):
        assert c[i] == min(k, total_rows)
        msg = exact.check_topk_parity(r[i, :c[i]].numpy(), d[i, :c[i]].numpy(), L[i], D[i])
        assert msg is None, msg
    # range search: concatenation of the shards' hit lists == range search over the whole matrix
    radius = float(np.sort(np.asarray(D[0], np.float64))[-1]) if total_rows >= 1 else 0.5
    got = idx.range_search(Q, radius)
    LR, DR = exact.range_search(full, Q, radius, space)
    for i in range(4):
        assert np.array_equal(got[i][1], np.asarray(LR[i], np.int64)), "sharded range rows differ"
        assert np.array_equal(got[i][0], np.asarray(DR[i], np.float32))
    dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
