"""``ShardedGpuIndex``: the reference ``Index`` protocol over the GPUs of one box, ONE PROCESS PER GPU.

Same constructor, protocol methods (``add`` / ``remove`` / ``search`` / ``rebuild``), ``is_rebuild_required``, ``_space``
and ``SearchResult`` as the reference ``Index`` (``src/mlvectordb/implementations/index.py:17-165``) and ``GpuIndex``;
it is what ``QueryProcessor(storage, index)`` (reference ``api/server.py:54``) gets when the server runs as one rank per
GPU (``torchrun``), and what ``bench.py`` times end to end at N > 1.  The calls are SPMD: every rank makes the same
calls with the same arguments and gets the same answers.

A namespace's rows are spread over the ranks (new blocks are cut so the ranks' live counts stay level, like
``MultiGpuIndex``); rank r's rows live in its own ``DeviceShard`` and are numbered ``r << 40 | local row``.  A search is
one fused kernel per rank -- local scan, peer-memory exchange of the k candidates over NVLink, merge
(``csrc/exchange.cuh``) -- for k <= 55 and small batches, the local tensor-core / scan path + NCCL all-gather +
``mlv_merge_topk`` otherwise (``sharded.ShardedIndex``).  Host state (row -> UUID per part, tombstone mirrors, counters)
is replicated on every rank: all ranks see every ``add`` / ``remove``, so no id ever crosses ranks at search time.
Ties are ordered by (distance, rank, local row); hnswlib's tie order is unspecified.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Mapping, Optional, Sequence
from uuid import UUID

import numpy as np

from .columns import ColumnCodec
from .index import PendingResults, _random_uuid_bytes, uuid_from_bytes
from .interfaces import SearchResult, VectorDTO, VectorProtocol
from .multi import PART_MASK, PART_SHIFT, split_block
from .shard import canonical_space


class _Part:
    """Replicated host view of one rank's rows of a namespace: row -> uuid bytes, tombstone mirror."""

    def __init__(self):
        self.ids = np.empty((16, 16), dtype=np.uint8)
        self.gone = np.zeros(16, dtype=bool)
        self.n = 0           # rows stored incl. tombstoned
        self.deleted = 0

    @property
    def live(self) -> int:
        return self.n - self.deleted

    def append(self, ids: np.ndarray) -> int:
        n = ids.shape[0]
        if self.n + n > self.ids.shape[0]:
            cap = max(self.n + n, 2 * self.ids.shape[0])
            grown = np.empty((cap, 16), dtype=np.uint8)
            grown[: self.n] = self.ids[: self.n]
            self.ids = grown
            gone = np.zeros(cap, dtype=bool)
            gone[: self.n] = self.gone[: self.n]
            self.gone = gone
        first = self.n
        self.ids[first:first + n] = ids
        self.gone[first:first + n] = False
        self.n += n
        return first

    def compact(self) -> None:
        keep = ~self.gone[: self.n]
        live = int(keep.sum())
        self.ids[:live] = self.ids[: self.n][keep]
        self.gone[: self.n] = False
        self.n, self.deleted = live, 0


class _ShardedNamespace:
    def __init__(self, dim: int, space: str, searcher, world: int):
        self.dim, self.space, self.searcher = dim, space, searcher
        self.parts = [_Part() for _ in range(world)]
        self.total = 0                  # reference index.py:27 _total_counts
        self.deleted = 0                # reference index.py:28 _deleted_counts
        self.rebuild_required = False
        self.lookup: Optional[Dict[bytes, int]] = {}   # uuid bytes -> global row; None = rebuild lazily
        self.codec = ColumnCodec()
        self.where_cache: Dict[tuple, object] = {}

    def touch(self) -> None:
        for f in self.where_cache.values():
            f.close()
        self.where_cache.clear()

    def uuid_of(self, global_row: int) -> UUID:
        return uuid_from_bytes(self.parts[global_row >> PART_SHIFT].ids[global_row & PART_MASK].tobytes())

    def lookup_table(self) -> Dict[bytes, int]:
        if self.lookup is None:
            table: Dict[bytes, int] = {}
            for p, part in enumerate(self.parts):
                raw = part.ids[: part.n].tobytes()
                for r in np.flatnonzero(~part.gone[: part.n]).tolist():
                    table[raw[16 * r: 16 * r + 16]] = (p << PART_SHIFT) | r
            self.lookup = table
        return self.lookup


class ShardedGpuIndex:
    def __init__(self, space: str = "l2", ef_construction: int = 200, M: int = 16, rebuild_threshold: float = 0.2,
                 device=None, group=None, capacity: int = 0, auto_compact: bool = True,
                 searcher_factory: Optional[Callable] = None):
        # ef_construction / M: signature compatibility (reference index.py:18); an exact scan has no graph parameters
        import torch.distributed as dist
        canonical_space(space)
        self._space = space
        self._rebuild_threshold = float(rebuild_threshold)
        self._auto_compact = bool(auto_compact)
        self._device, self._group = device, group
        self._capacity_hint = int(capacity)
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._ns: Dict[str, _ShardedNamespace] = {}
        # the device side is injectable so the host logic runs under gloo on a box without GPUs
        self._searcher_factory = searcher_factory or self._device_searcher

    def _device_searcher(self, dim: int, space: str, capacity: int):
        from .sharded import ShardedIndex
        bases = [r << PART_SHIFT for r in range(self.world)]
        return ShardedIndex(dim, space, 0, device=self._device, group=self._group, row_bases=bases,
                            capacity=-(-capacity // self.world) if capacity else 0)

    # ------------------------------------------------------------------ internals
    def _get_or_create(self, namespace: str, dim: int, metric: str, capacity: int = 0) -> _ShardedNamespace:
        ns = self._ns.get(namespace)
        if ns is None:
            space = canonical_space(metric)
            ns = _ShardedNamespace(dim, space, self._searcher_factory(dim, space, capacity or self._capacity_hint), self.world)
            self._ns[namespace] = ns
        return ns

    def _append(self, ns: _ShardedNamespace, data: np.ndarray, ids: np.ndarray, metadata=None) -> np.ndarray:
        """Cut the block over the ranks (fewest live rows first); this rank uploads its slice only, every rank records
        every slice's ids.  Returns the global rows of the block."""
        n = data.shape[0]
        sizes = split_block(n, [p.live for p in ns.parts])
        codes = ns.codec.encode_rows(metadata) if metadata is not None and any(metadata) else {}
        rows = np.empty(n, dtype=np.int64)
        at = 0
        for p, take in enumerate(sizes):
            if not take:
                continue
            first = ns.parts[p].append(ids[at:at + take])
            rows[at:at + take] = np.arange(first, first + take, dtype=np.int64) + (p << PART_SHIFT)
            if p == self.rank:
                got = ns.searcher.shard.add(data[at:at + take])
                assert got == first, "host id table out of step with the device matrix"
                for column, values in codes.items():
                    ns.searcher.shard.set_column(column, values[at:at + take], first)
            at += take
        ns.total += n
        ns.touch()
        return rows

    def _compact(self, ns: _ShardedNamespace) -> None:
        if ns.parts[self.rank].deleted:
            ns.searcher.shard.compact()              # survivors keep their order: every rank can renumber every part
        for part in ns.parts:
            part.compact()
        ns.total -= ns.deleted
        ns.deleted = 0
        ns.rebuild_required = False
        ns.lookup = None
        ns.touch()

    def _maybe_compact(self, ns: _ShardedNamespace) -> None:
        if ns.deleted / max(1, ns.total) < self._rebuild_threshold:
            return
        if not self._auto_compact:
            ns.rebuild_required = True           # reference index.py:86-89
            return
        self._compact(ns)

    def _local_filter(self, ns: _ShardedNamespace, filt):
        """Metadata constraints -> this rank's prepared filter over its own rows (``where_kernel``); None passes."""
        if filt is None:
            return None
        if not isinstance(filt, Mapping):
            raise ValueError("ShardedGpuIndex filters are metadata constraints {key: value | (op, a[, b])}")
        key = tuple(sorted(filt.items(), key=lambda kv: kv[0]))
        cached = ns.where_cache.get(key)
        if cached is None:
            self._order_columns(ns, filt)
            preds = ns.codec.predicates(filt)
            if preds is None:
                raise ValueError(f"constraints {dict(filt)!r} cannot be evaluated on the device columns")
            cached = ns.where_cache[key] = ns.searcher.shard.where(preds)
        return cached

    def _order_columns(self, ns: _ShardedNamespace, constraints: Mapping) -> None:
        """As ``GpuIndex._order_columns``: the codec is replicated, so every rank computes the same re-coding and applies
        it to its own rows."""
        from . import _capi
        for name in ns.codec.unordered_columns(constraints):
            perm = ns.codec.reorder(name)
            n_local = ns.parts[self.rank].n
            if perm is None or n_local == 0:
                continue
            column = ns.codec.column_index(name)
            codes = ns.searcher.shard.get_column(column, 0, n_local)
            has = codes != _capi.COLUMN_MISSING
            codes[has] = perm[codes[has]]
            ns.searcher.shard.set_column(column, codes, 0)

    def _results(self, ns: _ShardedNamespace, dists, rows, metric: str) -> List[SearchResult]:
        out = []
        for row, dist in zip(rows.tolist(), dists.tolist()):
            score = float(dist)
            if metric == "cosine":
                score = 1 - score                   # reference index.py:126-127
            out.append(SearchResult(vector_id=ns.uuid_of(int(row)), score=score))
        return out

    # ------------------------------------------------------------------ IndexProtocol
    def add(self, vectors: Iterable[VectorProtocol], namespace: str) -> None:
        """reference index.py:50-67"""
        vectors = list(vectors)
        if not vectors:
            return
        dim = vectors[0].values.shape[0]
        ns = self._get_or_create(namespace, dim, self._space)
        data = np.array([v.values for v in vectors], dtype=np.float32)
        if data.ndim != 2 or data.shape[1] != ns.dim:
            raise RuntimeError("Wrong dimensionality of the vectors")      # hnswlib's add_items error
        ids = np.frombuffer(b"".join(v.id.bytes for v in vectors), dtype=np.uint8).reshape(-1, 16)
        rows = self._append(ns, data, ids, [getattr(v, "metadata", None) for v in vectors])
        if ns.lookup is not None:
            for v, r in zip(vectors, rows.tolist()):
                ns.lookup[v.id.bytes] = r

    def remove(self, ids: Sequence[UUID], namespace: str) -> None:
        """reference index.py:69-89: unknown ids are ignored; the owning rank tombstones the row on its device."""
        ns = self._ns.get(namespace)
        if ns is None:
            return
        table = ns.lookup_table()
        mine, n = [], 0
        for uid in ids:
            g = table.pop(uid.bytes, None)
            if g is None:
                continue
            p, r = g >> PART_SHIFT, g & PART_MASK
            ns.parts[p].gone[r] = True
            ns.parts[p].deleted += 1
            n += 1
            if p == self.rank:
                mine.append(r)
        if mine:
            changed = ns.searcher.shard.mark_deleted(np.asarray(mine, dtype=np.uint64))
            assert changed == len(mine), "device tombstones out of step with the host id table"
        if n:
            ns.touch()
        ns.deleted += n
        self._maybe_compact(ns)

    def _prepare(self, query: VectorDTO, top_k: int, namespace: str):
        ns = self._ns.get(namespace)
        if ns is None:
            return None, None, 0
        k = min(int(top_k), ns.total - ns.deleted)
        q = np.asarray(query.values, dtype=np.float32).reshape(-1)
        if k < 1 or q.shape[0] != ns.dim:
            return None, None, 0     # the reference returns [] for all of these (index.py:98-119)
        return ns, q[None, :], k

    def search(self, query: VectorDTO, top_k: int, namespace: str, metric: str, filter=None) -> List[SearchResult]:  # noqa: A002
        """reference index.py:91-129 (``filter``: metadata constraints, additive).  Collective."""
        ns, q, k = self._prepare(query, top_k, namespace)
        if ns is None:
            return []
        dists, rows, counts = ns.searcher.search(q, k, filt=self._local_filter(ns, filter))
        c = int(counts[0])
        return self._results(ns, dists[0, :c], rows[0, :c], metric)

    def search_async(self, query: VectorDTO, top_k: int, namespace: str, metric: str) -> PendingResults:
        """``search`` that returns at once (collective: every rank submits the same sequence); ``.result()`` gives the
        ``List[SearchResult]``.  At most four in flight."""
        ns, q, k = self._prepare(query, top_k, namespace)
        if ns is None:
            return PendingResults(None, None, metric)
        return PendingResults(ns.searcher.search_async(q, k), ns, metric)

    def rebuild(self, source: Mapping[str, Iterable[VectorProtocol]], metric: str) -> None:
        """reference index.py:131-162: drop everything, re-add ``source`` with ``space=metric``."""
        self.close()
        for namespace, vectors in source.items():
            vectors = list(vectors)
            if not vectors:
                continue
            dim = vectors[0].values.shape[0]
            ns = self._get_or_create(namespace, dim, metric, capacity=len(vectors))
            data = np.array([v.values for v in vectors], dtype=np.float32)
            ids = np.frombuffer(b"".join(v.id.bytes for v in vectors), dtype=np.uint8).reshape(-1, 16)
            rows = self._append(ns, data, ids, [getattr(v, "metadata", None) for v in vectors])
            ns.lookup = {v.id.bytes: r for v, r in zip(vectors, rows.tolist())}

    def is_rebuild_required(self, namespace: str) -> bool:
        """reference index.py:164-165"""
        ns = self._ns.get(namespace)
        return bool(ns.rebuild_required) if ns is not None else False

    # ------------------------------------------------------------------ additive surface (as GpuIndex)
    def dimension(self, namespace: str) -> Optional[int]:
        ns = self._ns.get(namespace)
        return ns.dim if ns is not None else None

    def add_matrix(self, matrix: np.ndarray, namespace: str, ids: Optional[Sequence[UUID]] = None,
                   metadata: Optional[Sequence[Optional[Mapping]]] = None) -> np.ndarray:
        """Bulk ingest.  ``ids`` (or the generated ones) must be the same on every rank: pass them, or seed
        nothing and let rank 0's be broadcast.  Returns the rows' UUID bytes [n, 16]."""
        data = np.ascontiguousarray(matrix, dtype=np.float32)
        if data.ndim != 2:
            raise ValueError("matrix must be [n, dim]")
        if data.shape[0] == 0:
            return np.empty((0, 16), dtype=np.uint8)
        ns = self._get_or_create(namespace, data.shape[1], self._space, capacity=data.shape[0])
        if data.shape[1] != ns.dim:
            raise RuntimeError("Wrong dimensionality of the vectors")
        id_bytes = self._shared_ids(data.shape[0], ids)
        if metadata is not None and len(metadata) != data.shape[0]:
            raise ValueError("len(metadata) != rows")
        rows = self._append(ns, data, id_bytes, metadata)
        if data.shape[0] > 100_000:
            ns.lookup = None
        elif ns.lookup is not None:
            raw = id_bytes.tobytes()
            for i, r in enumerate(rows.tolist()):
                ns.lookup[raw[16 * i: 16 * i + 16]] = r
        return id_bytes

    def _shared_ids(self, n: int, ids) -> np.ndarray:
        if ids is not None:
            id_bytes = np.frombuffer(b"".join(u.bytes for u in ids), dtype=np.uint8).reshape(-1, 16)
            if id_bytes.shape[0] != n:
                raise ValueError("len(ids) != rows")
            return id_bytes
        import torch
        import torch.distributed as dist
        id_bytes = _random_uuid_bytes(n)
        if self.world > 1:                      # every rank must hold the same table: rank 0's ids win
            t = torch.from_numpy(id_bytes)
            if self._device is not None:
                t = t.to(self._device)
            dist.broadcast(t, dist.get_global_rank(self._group, 0) if self._group is not None else 0, group=self._group)
            id_bytes = t.cpu().numpy()
        return id_bytes

    def add_synthetic(self, namespace: str, n: int, dim: int, seed: int, scaled: bool = False) -> None:
        """Benchmark / parity input: generator rows 0..n-1 produced on the devices, contiguous blocks per rank (generator
        row g lives at rank g // ceil(n/world), local row g % ceil(n/world) of a fresh namespace)."""
        from .sharded import shard_range
        ns = self._get_or_create(namespace, dim, self._space, capacity=n)
        if dim != ns.dim:
            raise RuntimeError("Wrong dimensionality of the vectors")
        if ns.total:
            raise RuntimeError("add_synthetic needs a fresh namespace")
        ids = self._shared_ids(n, None)
        for p in range(self.world):
            lo, hi = shard_range(n, p, self.world)
            if hi > lo:
                first = ns.parts[p].append(ids[lo:hi])
                if p == self.rank:
                    assert ns.searcher.shard.add_synthetic(seed, lo, hi - lo, scaled) == first
        ns.total += n
        ns.lookup = None
        ns.touch()

    def search_batch(self, queries: np.ndarray, top_k: int, namespace: str, metric: Optional[str] = None, filter=None):  # noqa: A002
        """-> (global rows i64 [nq,k] (-1 padded; ``uuids_of`` decodes them), scores f32 [nq,k], counts i32 [nq])."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        nq = q.shape[0]
        ns = self._ns.get(namespace)
        k = min(int(top_k), (ns.total - ns.deleted) if ns is not None else 0)
        if ns is None or k < 1 or q.shape[1] != ns.dim:
            return (np.full((nq, 0), -1, np.int64), np.empty((nq, 0), np.float32), np.zeros(nq, np.int32))
        dists, rows, counts = ns.searcher.search(q, k, filt=self._local_filter(ns, filter))
        if (metric if metric is not None else self._space) == "cosine":
            dists = (1.0 - dists.astype(np.float64)).astype(np.float32)
        return rows, dists, counts

    def uuids_of(self, namespace: str, rows: np.ndarray) -> List[Optional[UUID]]:
        ns = self._ns[namespace]
        return [ns.uuid_of(int(r)) if r >= 0 else None for r in np.asarray(rows).reshape(-1)]

    def range_search(self, query: VectorDTO, radius: float, namespace: str, metric: str, filter=None) -> List[SearchResult]:  # noqa: A002
        """Every live row of every rank with hnswlib-form distance <= radius, nearest first.  Collective."""
        ns = self._ns.get(namespace)
        if ns is None or ns.total - ns.deleted == 0:
            return []
        q = np.asarray(query.values, dtype=np.float32).reshape(-1)
        if q.shape[0] != ns.dim:
            return []
        (dists, rows), = ns.searcher.range_search(q[None, :], float(radius), filt=self._local_filter(ns, filter))
        return self._results(ns, dists, rows, metric)

    def info(self, namespace: str) -> dict:
        ns = self._ns[namespace]
        return {"rows": sum(p.n for p in ns.parts), "live": ns.total - ns.deleted, "dim": ns.dim, "space": ns.space,
                "tombstones": sum(p.deleted for p in ns.parts), "rows_per_rank": [p.n for p in ns.parts], "world": self.world}

    def namespaces(self) -> List[str]:
        return list(self._ns)

    def close(self) -> None:
        for ns in self._ns.values():
            ns.touch()
            ns.searcher.close()
        self._ns.clear()
