"""Row-sharded exact search across the GPUs of one box.

One process per GPU (``torch.distributed``, NCCL over NVLink/NVSwitch).  Rank r owns the
contiguous row block ``[r * ceil(N/G), min(N, (r+1) * ceil(N/G)))`` of the namespace as one
``DeviceShard`` whose ``row_base`` is the block start, so every local result already carries
global rows.  A search is:

    local scan + top-k on every rank  ->  all-gather of the k candidates per query
    ((fp32 distance, int64 row) pairs, ``G * nq * k * 12`` bytes)  ->  ``mlv_merge_topk`` on every rank

The exchange is the only collective on the path; it is latency-bound (120 bytes per rank for a
batch-1, k=10 query), so it is issued as two ``all_gather_into_tensor`` calls on the stream the
scan ran on.  The reference has no counterpart (single process, README sketch only:
``README.md:142-155``); the identity it relies on -- top-k of a union of row shards equals the
merge of the shards' top-k -- is checked in ``tests/test_gpu_parity.py``.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import _capi


def shard_range(total_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous row block of ``rank``: ceil(N/G) rows each, the tail ranks may be short or empty."""
    per = -(-total_rows // world)
    lo = min(total_rows, rank * per)
    hi = min(total_rows, lo + per)
    return lo, hi


def gather_candidates(dists: torch.Tensor, rows: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather every rank's ``[nq, k]`` candidates -> ``[G, nq, k]`` (rank-major = ascending row_base).

    Works on any backend (NCCL on the GPUs, gloo in the CPU tests): the layout is what
    ``mlv_merge_topk`` consumes.
    """
    world = dist.get_world_size(group)
    nq, k = dists.shape
    # concatenation along dim 0 is the one output form every backend accepts; viewed as [G, nq, k]
    gd = torch.empty((world * nq, k), dtype=dists.dtype, device=dists.device)
    gr = torch.empty((world * nq, k), dtype=rows.dtype, device=rows.device)
    dist.all_gather_into_tensor(gd, dists.contiguous(), group=group)
    dist.all_gather_into_tensor(gr, rows.contiguous(), group=group)
    return gd.view(world, nq, k), gr.view(world, nq, k)


def merge_topk_device(gd: torch.Tensor, gr: torch.Tensor, k: int):
    """``mlv_merge_topk`` on the current CUDA stream; inputs ``[G, nq, k]`` device tensors."""
    G, nq, kk = gd.shape
    assert kk == k and gd.is_cuda and gd.dtype == torch.float32 and gr.dtype == torch.int64
    out_d = torch.empty((nq, k), dtype=torch.float32, device=gd.device)
    out_r = torch.empty((nq, k), dtype=torch.int64, device=gd.device)
    out_c = torch.empty((nq,), dtype=torch.int32, device=gd.device)
    stream = torch.cuda.current_stream(gd.device).cuda_stream
    st = _capi.lib().mlv_merge_topk(gd.device.index, C.c_void_p(gd.data_ptr()), C.c_void_p(gr.data_ptr()), G, nq, k,
                                    C.c_void_p(out_d.data_ptr()), C.c_void_p(out_r.data_ptr()),
                                    C.c_void_p(out_c.data_ptr()), C.c_void_p(stream))
    _capi.check(st)
    return out_d, out_r, out_c


class _Done:
    def __init__(self, out):
        self._out = out

    def result(self):
        return self._out


class ShardedIndex:
    """One namespace, rows sharded over the ranks of ``group``.  Every rank calls every method."""

    # batches of this many queries or more (mlv gemm_min_nq) run the local tensor-core path (csrc/gemm_kernel.cuh) and
    # merge through NCCL; smaller ones use the fused scan + peer-memory exchange kernel
    EXCHANGE_MAX_NQ = 9

    def __init__(self, dim: int, space: str, total_rows: int, device: Optional[torch.device] = None, group=None,
                 local_search: Optional[Callable] = None, merge: Optional[Callable] = None, fused_exchange: bool = True,
                 local_range: Optional[Callable] = None, order_hits: Optional[Callable] = None,
                 row_bases: Optional[Sequence[int]] = None, capacity: int = 0, shard=None):
        """``row_bases`` (one per rank) switches from fixed contiguous blocks of ``total_rows`` to PART mode: rank r's
        rows are numbered ``row_bases[r] + local row`` and the shard grows by appends (``ShardedGpuIndex``); ties are
        then ordered by (distance, rank, local row)."""
        self.dim, self.space, self.total_rows = int(dim), space, int(total_rows)
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.row_bases = [int(b) for b in row_bases] if row_bases is not None else None
        if self.row_bases is not None:
            assert len(self.row_bases) == self.world
            self.lo, self.hi = self.row_bases[self.rank], None
        else:
            self.lo, self.hi = shard_range(self.total_rows, self.rank, self.world)
        self.device = device
        self.shard = shard          # injected stand-in (CPU tests); the real DeviceShard is created below
        # the two device steps are injectable so the collective plumbing can be exercised on CPU (gloo)
        self._local_search = local_search or self._device_local_search
        self._merge = merge or merge_topk_device
        self._local_range = local_range or self._device_local_range
        self._order_hits = order_hits or self._device_order_hits
        self.merge_launches = 0
        self.exchange = None
        if local_search is None:
            from .shard import DeviceShard
            cap = int(capacity) if self.hi is None else max(self.hi - self.lo, 1)
            self.shard = DeviceShard(dim, space, capacity=cap, device=device.index, row_base=self.lo)
            if self.world > 1 and fused_exchange:
                self._setup_exchange()

    def _setup_exchange(self) -> None:
        """Swap CUDA IPC handles of the per-rank exchange buffers (``csrc/exchange.cuh``) and attach them.

        If any rank cannot map its peers (IPC disabled in the container), every rank stays on the
        NCCL all-gather path -- still the GPU path, just two collectives and a merge launch more.
        """
        from .shard import Exchange
        ok, ex = 1, None
        try:
            ex = Exchange(self.device.index, self.world, self.rank)
        except RuntimeError:
            ok = 0
        handles = [None] * self.world
        dist.all_gather_object(handles, (ok, ex.handle if ex else b""), group=self.group)
        if all(h[0] for h in handles):
            try:
                ex.connect([h[1] for h in handles])
            except RuntimeError:
                ok = 0
        else:
            ok = 0
        flags = [None] * self.world
        dist.all_gather_object(flags, ok, group=self.group)
        if all(flags):
            bases = self.row_bases or [shard_range(self.total_rows, r, self.world)[0] for r in range(self.world)]
            self.shard.attach_exchange(ex, bases)
            self.exchange = ex
        elif ex is not None:
            ex.close()

    # -- data -------------------------------------------------------------------------------
    def add_synthetic(self, seed: int, scaled: bool) -> None:
        if self.hi > self.lo:
            self.shard.add_synthetic(seed, self.lo, self.hi - self.lo, scaled)

    def add_rows(self, local_rows: np.ndarray) -> None:
        assert local_rows.shape[0] == self.hi - self.lo
        if self.hi > self.lo:
            self.shard.add(local_rows)

    # -- metadata columns / filters: purely local, a row's values live on the rank that owns the row ------
    def set_column(self, column: int, local_values) -> None:
        """int32 codes of this rank's row block for ``column`` (``DeviceShard.set_column``)."""
        assert len(local_values) == self.hi - self.lo
        if self.hi > self.lo:
            self.shard.set_column(column, local_values)

    def where(self, predicates):
        """Collective only in the sense that every rank prepares the same predicate over its own rows; pass the
        result as ``filt=`` to ``search`` / ``range_search`` on that rank.  No exchange: the filter restricts the
        local scan, the candidates merge as before."""
        return self.shard.where(predicates)

    def _bound(self, filt):
        from .shard import _bound
        return _bound(self.shard, filt)

    # -- search -----------------------------------------------------------------------------
    def _device_local_search(self, q: torch.Tensor, k: int):
        nq = q.shape[0]
        d = torch.empty((nq, k), dtype=torch.float32, device=q.device)
        r = torch.empty((nq, k), dtype=torch.int64, device=q.device)
        c = torch.empty((nq,), dtype=torch.int32, device=q.device)
        stream = torch.cuda.current_stream(q.device).cuda_stream
        self.shard.search_device(q.data_ptr(), nq, k, d.data_ptr(), r.data_ptr(), c.data_ptr(), stream=stream)
        return d, r, c

    def search_device(self, q: torch.Tensor, k: int, filt=None):
        """``q``: [nq, dim] fp32 tensor on this rank's device (same on every rank).  Returns the
        global top-k ``(dists [nq,k], rows [nq,k], counts [nq])`` on every rank, nothing synchronised.
        ``filt``: this rank's prepared filter (``where``), every rank passing its own."""
        if filt is not None:
            with self._bound(filt):
                return self.search_device(q, k)
        if self.exchange is not None and q.shape[0] < self.EXCHANGE_MAX_NQ and self.shard.exchange_supported(k):
            # one kernel per group of <= 8 queries: scan + peer-memory exchange + merge (csrc/exchange.cuh)
            nq = q.shape[0]
            d = torch.empty((nq, k), dtype=torch.float32, device=q.device)
            r = torch.empty((nq, k), dtype=torch.int64, device=q.device)
            c = torch.empty((nq,), dtype=torch.int32, device=q.device)
            stream = torch.cuda.current_stream(q.device).cuda_stream
            self.shard.search_exchange_device(q.data_ptr(), nq, k, d.data_ptr(), r.data_ptr(), c.data_ptr(), stream=stream)
            return d, r, c
        d, r, c = self._local_search(q, k)
        if self.world == 1:
            return d, r, c
        gd, gr = gather_candidates(d, r, self.group)
        self.merge_launches += 1
        return self._merge(gd, gr, k)

    def search(self, queries: np.ndarray, k: int, filt=None):
        """Host buffers in, host buffers out (pinned staging, H2D + D2H inside the call)."""
        if filt is not None:
            with self._bound(filt):
                return self.search(queries, k)
        q = np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, self.dim)
        nq = q.shape[0]
        if self.exchange is not None and nq < self.EXCHANGE_MAX_NQ and self.shard.exchange_supported(k):
            # straight through the C host entry point: one H2D, one kernel per query group, one D2H, one sync
            out = self.shard.search_exchange(q, k)
            if (out[2] < 0).any():
                raise RuntimeError("sharded search: a peer rank did not post its candidates within the exchange timeout")
            return out
        if self.device is None:      # CPU plumbing tests (gloo): nothing to stage
            d, r, c = self.search_device(torch.from_numpy(q), k)
            return d.numpy(), r.numpy(), c.numpy()
        st = self._staging(nq, k)
        st["q"][:nq].copy_(torch.from_numpy(q))
        qd = st["q"][:nq].to(self.device, non_blocking=True)
        d, r, c = self.search_device(qd, k)
        st["d"][:nq].copy_(d, non_blocking=True)
        st["r"][:nq].copy_(r, non_blocking=True)
        st["c"][:nq].copy_(c, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()   # the one host sync of the call
        out = (st["d"][:nq].numpy().copy(), st["r"][:nq].numpy().copy(), st["c"][:nq].numpy().copy())
        if (out[2] < 0).any():
            raise RuntimeError("sharded search: a peer rank did not post its candidates within the exchange timeout")
        return out

    def search_async(self, queries: np.ndarray, k: int):
        """Collective.  Start a search and return a handle whose ``.result()`` gives what ``search`` returns;
        a caller that keeps a few in flight (at most four) hides each request's copies and launch latency behind the others' scans.
        Needs the fused exchange path (k small, batch < EXCHANGE_MAX_NQ); otherwise the search runs synchronously."""
        q = np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, self.dim)
        if self.device is not None and self.shard is not None and q.shape[0] < self.EXCHANGE_MAX_NQ and (
                self.world == 1 or (self.exchange is not None and self.shard.exchange_supported(k))):
            return self.shard.submit(q, k, exchange=self.world > 1)
        return _Done(self.search(q, k))

    # -- range search ------------------------------------------------------------------------
    def _local_empty(self) -> bool:
        return (self.shard.rows == 0) if self.hi is None else (self.hi == self.lo)

    def _device_local_range(self, queries: np.ndarray, radius: float):
        if self._local_empty():
            return [(np.empty(0, np.float32), np.empty(0, np.int64)) for _ in range(queries.shape[0])]
        return self.shard.range_search(queries, radius)   # rows already carry row_base

    def range_search(self, queries: np.ndarray, radius: float, filt=None):
        """Collective.  Every live row of every shard with hnswlib-form distance <= radius, per query
        ascending (distance, global row): the concatenation of the shards' hit lists.
        The exchange is one all-gather of the hit counts and one of the lists padded to the longest."""
        if filt is not None:
            with self._bound(filt):
                return self.range_search(queries, radius)
        q = np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, self.dim)
        nq = q.shape[0]
        if self.exchange is not None and self.shard.range_exchange_supported():
            # one fused kernel per query and rank: local range scan, the last CTA sorts the hits, exchanges them over
            # peer memory and merges (csrc/exchange.cuh); the lists come back in mapped host memory, one sync per call
            fused = self.shard.range_search_exchange(q, float(radius))
            if fused is not None:
                return fused
            # some rank found more hits than its share of the exchange slot (every rank saw that): all-gather path below
        local = self._local_range(q, float(radius))
        if self.world == 1:
            return local
        dev = self.device if self.device is not None else torch.device("cpu")
        counts = torch.tensor([len(d) for d, _ in local], dtype=torch.int64, device=dev)
        all_counts = torch.empty((self.world * nq,), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(all_counts, counts, group=self.group)
        all_counts = all_counts.view(self.world, nq).cpu().numpy()
        m = int(all_counts.max())
        if m == 0:
            return [(np.empty(0, np.float32), np.empty(0, np.int64)) for _ in range(nq)]
        pd = np.full((nq, m), np.inf, np.float32)
        pr = np.full((nq, m), -1, np.int64)
        for i, (d, r) in enumerate(local):
            pd[i, :len(d)], pr[i, :len(r)] = d, r
        gd = torch.empty((self.world * nq, m), dtype=torch.float32, device=dev)
        gr = torch.empty((self.world * nq, m), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(gd, torch.from_numpy(pd).to(dev), group=self.group)
        dist.all_gather_into_tensor(gr, torch.from_numpy(pr).to(dev), group=self.group)
        # concatenation of the shards' lists per query, ordered (distance, global row) on the device
        gd, gr = gd.view(self.world, nq, m), gr.view(self.world, nq, m)
        totals = all_counts.sum(axis=0)
        out = []
        for i in range(nq):
            # keyed by position in the rank-major concatenation (global rows may not fit the key's 32 row bits): every
            # list is ascending (distance, row) and ranks ascend with row_base, so position order is row order
            rows_i = gr[:, i, :].reshape(-1).contiguous()
            pos = torch.where(rows_i >= 0, torch.arange(rows_i.numel(), dtype=torch.int64, device=rows_i.device), rows_i)
            d, p = self._order_hits(gd[:, i, :].reshape(-1).contiguous(), pos)
            n = int(totals[i])
            out.append((d[:n].cpu().numpy().copy(), rows_i[p[:n]].cpu().numpy().copy()))
        return out

    def _device_order_hits(self, d: torch.Tensor, r: torch.Tensor):
        """(distance, global row) pairs of one query -> ascending; padding (row -1) last.  ``mlv_index_order_pairs_device``."""
        od, orr = torch.empty_like(d), torch.empty_like(r)
        stream = torch.cuda.current_stream(d.device).cuda_stream
        self.shard.order_pairs_device(d.data_ptr(), r.data_ptr(), d.numel(), od.data_ptr(), orr.data_ptr(), stream=stream)
        return od, orr

    def _staging(self, nq: int, k: int):
        """Pinned host staging reused across calls (one allocation per (nq, k) growth)."""
        st = getattr(self, "_stage", None)
        if st is None or st["q"].shape[0] < nq or st["d"].shape[1] != k:
            cap = max(nq, 8)
            st = {"q": torch.empty((cap, self.dim), dtype=torch.float32).pin_memory(),
                  "d": torch.empty((cap, k), dtype=torch.float32).pin_memory(),
                  "r": torch.empty((cap, k), dtype=torch.int64).pin_memory(),
                  "c": torch.empty((cap,), dtype=torch.int32).pin_memory()}
            self._stage = st
        return st

    def close(self) -> None:
        if self.shard is not None:
            self.shard.close()
            self.shard = None
        if self.exchange is not None:
            if dist.is_initialized():
                dist.barrier(group=self.group)  # nobody unmaps a buffer a peer may still write
            self.exchange.close()
            self.exchange = None
