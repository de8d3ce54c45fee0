"""One rank of the multi-GPU test (launched by tests/test_gpu_sharded.py, one process per GPU).

usage: _nccl_worker.py RANK WORLD PORT
Every rank holds a row shard (``ShardedIndex``); rank 0 additionally holds the whole matrix in one
``DeviceShard`` and checks that the sharded results -- fused peer-memory exchange for k <= 55,
NCCL all-gather + merge kernel for larger k, tensor-core path for batches -- are bit-identical to
the unsharded search.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, port = (int(x) for x in sys.argv[1:4])
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    device = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    from mlvectordb_b200 import DeviceShard
    from mlvectordb_b200.sharded import ShardedIndex
    from oracle import synthetic

    def check(total_rows, dim, space, k, nq, expect_exchange, delete_every=0):
        idx = ShardedIndex(dim, space, total_rows, device=device)
        idx.add_synthetic(7, scaled=True)
        assert (idx.exchange is not None) == (world > 1), "peer-memory exchange was not set up"
        Q = synthetic.queries(9, nq, dim)
        if total_rows > 5:
            Q[0] = synthetic.rows(7, total_rows - 2, 1, dim, scaled=True)[0]   # exact match in the last shard
        if delete_every and idx.hi > idx.lo:
            dead = np.arange(idx.lo, idx.hi, delete_every, dtype=np.uint64) - idx.lo
            idx.shard.mark_deleted(dead)
        assert idx.shard.exchange_supported(k) == expect_exchange
        launches0 = idx.merge_launches
        for rep in range(3):                      # repeated: parity slots of the exchange get reused
            d, r, c = idx.search(Q, k)
        used_exchange = expect_exchange and nq < ShardedIndex.EXCHANGE_MAX_NQ
        assert (idx.merge_launches == launches0) == used_exchange
        if rank == 0:
            whole = DeviceShard(dim, space, capacity=total_rows, device=rank)
            whole.add_synthetic(7, 0, total_rows, True)
            if delete_every:
                dead_all = np.concatenate([np.arange(lo, hi, delete_every, dtype=np.uint64)
                                           for lo, hi in (idx_range(total_rows, rr, world) for rr in range(world)) if hi > lo])
                whole.mark_deleted(dead_all)
            wd, wr, wc = whole.search(Q, k)
            assert np.array_equal(c, wc), (c, wc)
            assert np.array_equal(r, wr), f"rows differ ({space}, n={total_rows}, k={k}, nq={nq})"
            assert np.array_equal(d, wd, equal_nan=True)
            if total_rows > 5 and space != "ip" and not delete_every:   # the planted exact match wins (not under ip)
                assert r[0, 0] == total_rows - 2
            if k == 10 and nq <= 8:
                radius = float(wd[0, min(9, wc[0] - 1)])            # <= 10 hits for query 0, more for none
                wr_all = whole.range_search(Q, radius)
            whole.close()
        if k == 10 and nq <= 8:
            rad = torch.zeros(1, dtype=torch.float64, device=device)
            if rank == 0:
                rad[0] = radius
            dist.broadcast(rad, 0)
            got = idx.range_search(Q, float(rad.item()))
            if rank == 0:
                for (gd, gr), (hd, hr) in zip(got, wr_all):
                    assert np.array_equal(gr, hr), "sharded range rows differ from unsharded"
                    assert np.array_equal(gd, hd)
                assert len(got[0][1]) >= 1
        idx.close()

    def check_single(total_rows, dim, space, k, tiers):
        """Single queries take a {first tier, conditional fp32} pair of launches on every rank (scan_kernel.cuh, shadow
        scan); ``tiers[rank % len(tiers)]`` is this rank's "scan_half": 1 = fp16 shadow first, -1 = too small here, so
        its first launch is the fp32 scan -- the ranks only have to agree on the number of launches."""
        idx = ShardedIndex(dim, space, total_rows, device=device)
        idx.add_synthetic(11, scaled=True)
        idx.shard.set_tuning("scan_half", tiers[rank % len(tiers)])
        nq = 7
        Q = synthetic.queries(12, nq, dim)
        if total_rows > 5:
            Q[0] = synthetic.rows(11, total_rows - 2, 1, dim, scaled=True)[0]
        got = [idx.search(Q[i:i + 1], k) for i in range(nq)]
        handles = [idx.search_async(Q[i:i + 1], k) for i in range(4)]      # four in flight: the exchange's limit
        got_async = [hnd.result() for hnd in handles]
        st = idx.shard.gemm_stats()
        if tiers[rank % len(tiers)] == 1 and idx.hi > idx.lo:
            assert st["half_scan_queries"] >= nq, st
        # range search with the same per-rank tiers (shadow range scan on some ranks, fp32 on others): radius = the 5th
        # distance of query 0, agreed through rank 0
        rad = torch.zeros(1, dtype=torch.float64, device=device)
        if rank == 0:
            rad[0] = float(got[0][0][0, min(4, int(got[0][2][0]) - 1)]) if int(got[0][2][0]) > 0 else 1.0
        dist.broadcast(rad, 0)
        got_range = idx.range_search(Q[:3], float(rad.item()))
        if rank == 0:
            whole = DeviceShard(dim, space, capacity=total_rows, device=rank)
            whole.add_synthetic(11, 0, total_rows, True)
            whole.set_tuning("scan_half", 0)
            for (gd, gr), (hd, hr) in zip(got_range, whole.range_search(Q[:3], float(rad.item()))):
                assert np.array_equal(gr, hr) and np.array_equal(gd, hd), f"sharded range (tiers={tiers}) differs from unsharded"
            for i in range(nq):
                ref = whole.search(Q[i:i + 1], k)
                for a, b in zip(got[i], ref):
                    assert np.array_equal(a, b, equal_nan=True), f"single query {i} differs ({space}, tiers={tiers})"
            for i in range(4):
                ref = whole.search(Q[i:i + 1], k)
                for a, b in zip(got_async[i], ref):
                    assert np.array_equal(a, b, equal_nan=True), f"async single query {i} differs"
            whole.close()
        idx.close()

    def check_filtered(total_rows, dim, space, k, nq):
        """metadata column sharded with the rows, predicate evaluated per rank, filtered sharded search ==
        filtered unsharded search (fused exchange and NCCL paths)"""
        idx = ShardedIndex(dim, space, total_rows, device=device)
        idx.add_synthetic(11, scaled=True)
        buckets = synthetic.buckets(11, 0, total_rows)
        idx.set_column(0, buckets[idx.lo:idx.hi])
        Q = synthetic.queries(12, nq, dim)
        preds = [(0, "<", 7)]
        f = idx.where(preds) if idx.hi > idx.lo else None
        d, r, c = idx.search(Q, k, filt=f)
        radius = float(d[0, c[0] - 1])                      # every rank holds the same global result
        hits = idx.range_search(Q[:2], radius, filt=f)
        if rank == 0:
            whole = DeviceShard(dim, space, capacity=total_rows, device=rank)
            whole.add_synthetic(11, 0, total_rows, True)
            whole.set_column(0, buckets)
            wf = whole.where(preds)
            wd, wr, wc = whole.search(Q, k, wf)
            assert np.array_equal(c, wc) and np.array_equal(r, wr) and np.array_equal(d, wd, equal_nan=True)
            assert (buckets[r[c[:, None] > np.arange(k)[None, :]]] < 7).all()
            for (gd, gr), (hd, hr) in zip(hits, whole.range_search(Q[:2], radius, wf)):
                assert np.array_equal(gr, hr) and np.array_equal(gd, hd)
            assert len(hits[0][1]) == c[0]
            whole.close()
        idx.close()

    def check_big_range(total_rows, dim):
        """hit lists far longer than one CTA sorts (> 8192 per query): the concatenated shard lists are ordered by the
        device-wide bitonic network (mlv_index_order_pairs_device) and equal the unsharded range search"""
        idx = ShardedIndex(dim, "l2", total_rows, device=device)
        idx.add_synthetic(13, scaled=False)
        Q = synthetic.queries(14, 2, dim)
        rad = torch.zeros(1, dtype=torch.float64, device=device)
        whole = None
        if rank == 0:
            whole = DeviceShard(dim, "l2", capacity=total_rows, device=rank)
            whole.add_synthetic(13, 0, total_rows, False)
            d1000 = whole.search(Q[:1], 1000)[0][0, 999]
            for mult in (1.2, 1.5, 2.0, 3.0, 5.0):
                if len(whole.range_search(Q[:1], float(d1000) * mult)[0][1]) > 12_000:
                    break
            rad[0] = float(d1000) * mult
        dist.broadcast(rad, 0)
        got = idx.range_search(Q, float(rad.item()))
        if rank == 0:
            want = whole.range_search(Q, float(rad.item()))
            assert len(want[0][1]) > 8192, len(want[0][1])
            for (gd, gr), (hd, hr) in zip(got, want):
                assert np.array_equal(gr, hr) and np.array_equal(gd, hd)
            whole.close()
        idx.close()

    def check_protocol_index():
        """``ShardedGpuIndex`` (reference Index protocol, one process per GPU): every rank makes the same calls; the
        answers equal a single ``GpuIndex`` holding all rows on rank 0 (ids; scores bit for bit)."""
        from uuid import UUID
        from mlvectordb_b200 import GpuIndex, VectorDTO
        from mlvectordb_b200.sharded_index import ShardedGpuIndex

        class V:
            def __init__(self, uid, values, metadata):
                self.id, self.values, self.metadata = uid, values, metadata

        n, dim = 40_000, 64
        X = synthetic.rows(21, 0, n, dim, scaled=True)
        vecs = [V(UUID(int=10_000 + i), X[i], {"b": i % 10}) for i in range(n)]
        sharded = ShardedGpuIndex(space="cosine", device=device)
        single = GpuIndex(space="cosine", device=rank) if rank == 0 else None
        for lo, hi in ((0, 5), (5, 10_000), (10_000, n)):
            sharded.add(vecs[lo:hi], "ns")
            if single:
                single.add(vecs[lo:hi], "ns")
        per_rank = sharded.info("ns")["rows_per_rank"]
        assert sum(per_rank) == n and max(per_rank) - min(per_rank) <= 1
        Q = synthetic.queries(22, 4, dim)
        Q[0] = X[n - 3]

        def compare(k, **kw):
            for q in Q:
                got = sharded.search(VectorDTO(values=q), k, "ns", "cosine", **kw)
                if single:
                    want = single.search(VectorDTO(values=q), k, "ns", "cosine", **kw)
                    assert [h.vector_id for h in got] == [h.vector_id for h in want], f"k={k} {kw}"
                    assert [h.score for h in got] == [h.score for h in want]

        compare(10)                               # fused peer-memory exchange
        compare(200)                              # NCCL all-gather + merge kernel
        compare(10, filter={"b": 7})              # per-rank device predicate
        pend = [sharded.search_async(VectorDTO(values=Q[i]), 10, "ns", "cosine") for i in range(2)]   # two in flight
        for i, p in enumerate(pend):
            got = p.result()
            if single:
                assert [h.vector_id for h in got] == [h.vector_id for h in single.search(VectorDTO(values=Q[i]), 10, "ns", "cosine")]
        gone = [v.id for v in vecs[1::3]]
        sharded.remove(gone, "ns")                # 33 % >= 0.2: every rank compacts its shard, id tables renumbered
        if single:
            single.remove(gone, "ns")
        assert sharded.info("ns")["tombstones"] == 0 and sharded.info("ns")["live"] == n - len(gone)
        compare(10)
        got = sharded.range_search(VectorDTO(values=Q[1]), 0.75, "ns", "cosine")
        if single:
            want = single.range_search(VectorDTO(values=Q[1]), 0.75, "ns", "cosine")
            assert [h.vector_id for h in got] == [h.vector_id for h in want] and len(want) > 0
        sharded.rebuild({"fresh": vecs[:100]}, "l2")
        hit = sharded.search(VectorDTO(values=X[42]), 1, "fresh", "l2")
        assert hit[0].vector_id == vecs[42].id and hit[0].score == 0.0 and sharded.namespaces() == ["fresh"]
        sharded.close()
        if single:
            single.close()

    def idx_range(n, r, w):
        from mlvectordb_b200.sharded import shard_range
        return shard_range(n, r, w)

    check(200_003, 64, "cosine", 10, 5, True)              # fused exchange, ragged shards
    check(200_003, 64, "l2", 10, 19, True)                 # several launch groups per call (nq > 8)
    check(50_000, 96, "ip", 13, 3, True, delete_every=3)   # tombstones
    check(1, 32, "l2", 4, 2, True)                         # rank 1.. hold nothing: exchange_only_kernel
    check(200_003, 64, "l2", 40, 3, True)                  # larger k: bitonic final select in the last CTA
    check(200_003, 64, "cosine", 100, 4, False)            # k too large for the fused exchange: NCCL path
    check(120_000, 128, "l2", 10, 300, True)               # large batch: local tensor-core path + NCCL merge
    check_big_range(120_001, 16)                           # > 8192 hits per query: device-wide ordering
    check_filtered(150_001, 64, "cosine", 10, 4)           # filtered, fused exchange
    check_filtered(150_001, 64, "l2", 100, 3)              # filtered, NCCL merge path
    check_single(200_003, 64, "cosine", 10, (1,))          # shadow scan on every rank + conditional fp32 launch
    check_single(200_003, 96, "l2", 16, (1, -1))           # mixed: some ranks offer the shadow, some the fp32 scan
    check_single(150_000, 64, "ip", 10, (-1,))             # nobody has a shadow: fp32 first launches, second ones skipped
    check_single(1, 32, "l2", 1, (1,))                     # ranks 1.. hold nothing: exchange_only_kernel pairs
    check_protocol_index()                                 # reference Index protocol over the ranks
    dist.barrier()
    print(f"rank {rank} ok", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
