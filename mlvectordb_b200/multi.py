"""``MultiGpuIndex``: the drop-in index over SEVERAL GPUs of ONE process.

The reference server is a single process (``src/mlvectordb/api/server.py:54-72``: one uvicorn worker, one ``Index``);
its README only sketches sharding (``README.md:142-155``).  ``ShardedIndex`` (``sharded.py``) is the one-process-per-GPU
form the benchmarks use; this class is the form that server can use unchanged:
``QueryProcessor(storage, MultiGpuIndex(space="cosine", devices=[0, 1, 2, 3]))``.  Same constructor, protocol methods,
``is_rebuild_required``, ``_space`` and ``SearchResult`` as ``GpuIndex`` / the reference ``Index``
(``implementations/index.py:17-165``).

Every device holds a ``GpuIndex`` part; a namespace's rows are spread over the parts (new rows go where fewest live
rows are, whole blocks at a time).  A search is the multi-GPU identity of SURVEY.md section 8e inside one process:

    every part scans its rows on its own device and stream (``mlv_index_search_device``, all devices at once)
    -> the k candidates of each part are copied to the first device over NVLink (peer copies)
    -> ``mlv_merge_topk`` (the final select kernel of the NCCL path) orders them by (distance, part, row)

so no distance is computed and no candidate is ordered on the host.  Removal, rebuild, compaction, metadata columns and
filters are per part (a row's id, tombstone and column values live with the row).
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Mapping, Optional, Sequence
from uuid import UUID

import numpy as np

from .index import GpuIndex
from .interfaces import SearchResult, VectorDTO, VectorProtocol
from .shard import _bound, canonical_space

PART_SHIFT = 40                      # global row = part << 40 | local row (a shard holds < 2^32 rows)
PART_MASK = (1 << PART_SHIFT) - 1


def split_block(n: int, live_counts: Sequence[int]) -> List[int]:
    """How many of ``n`` new rows each part takes so that the parts' live counts end up as level as possible
    (water filling: the emptiest parts are raised first); sums to ``n``."""
    counts = [int(c) for c in live_counts]
    if not counts:
        raise ValueError("no parts")
    n = int(n)
    lo, hi = min(counts), max(counts) + n
    while lo < hi:                                   # smallest level L with sum(max(0, L - c)) >= n
        mid = (lo + hi) // 2
        if sum(max(0, mid - c) for c in counts) >= n:
            hi = mid
        else:
            lo = mid + 1
    level = lo
    take = [max(0, level - 1 - c) for c in counts]   # everyone below L - 1 is raised to it ...
    left = n - sum(take)
    for i in sorted(range(len(counts)), key=lambda j: (counts[j] + take[j], j)):
        if left == 0:
            break
        if counts[i] + take[i] < level:              # ... and the remainder goes one row each to the lowest parts
            take[i] += 1
            left -= 1
    assert left == 0 and sum(take) == n
    return take


class PartwiseFilter:
    """A filter every part evaluates over its own rows (metadata constraints or ``callable(uuid)``): what
    ``MultiGpuIndex.where`` / ``prepare_filter`` hand to ``GpuQueryProcessor``, which passes it back as ``filter=``."""

    def __init__(self, spec):
        self.spec = spec

    def close(self) -> None:
        pass


class MultiGpuIndex:
    def __init__(self, space: str = "l2", ef_construction: int = 200, M: int = 16, rebuild_threshold: float = 0.2,
                 devices: Optional[Sequence[int]] = None, capacity: int = 0, auto_compact: bool = True,
                 fanout: Optional[Callable] = None, order_pairs: Optional[Callable] = None):
        canonical_space(space)
        if devices is None:
            from . import _capi
            devices = list(range(max(1, _capi.lib().mlv_device_count())))
        if not devices:
            raise ValueError("devices must name at least one GPU")
        self._space = space
        self._devices = [int(d) for d in devices]
        self._parts = [GpuIndex(space=space, ef_construction=ef_construction, M=M, rebuild_threshold=rebuild_threshold, device=d,
                                capacity=int(capacity) // len(self._devices), auto_compact=auto_compact) for d in self._devices]
        # the device step is injectable so the host logic runs on a box without GPUs (tests/test_host_logic_cpu.py)
        self._fanout = fanout or self._device_fanout
        self._order_pairs = order_pairs or self._device_order_pairs

    # ------------------------------------------------------------------ helpers
    def _live(self, part: GpuIndex, namespace: str) -> int:
        ns = part._ns.get(namespace)
        return (ns.total - ns.deleted) if ns is not None else 0

    def _holders(self, namespace: str):
        return [(i, p, p._ns[namespace]) for i, p in enumerate(self._parts) if namespace in p._ns]

    def _results(self, namespace: str, dists, rows, count: int, metric: str) -> List[SearchResult]:
        out = []
        for row, dist in zip(rows[:count].tolist(), dists[:count].tolist()):
            part, local = int(row) >> PART_SHIFT, int(row) & PART_MASK
            score = float(dist)
            if metric == "cosine":
                score = 1 - score                      # reference index.py:126-127
            out.append(SearchResult(vector_id=self._parts[part]._ns[namespace].uuid_of(local), score=score))
        return out

    def _part_filters(self, holders, filt):
        """Per-part prepared filters (or masks) for a filter given once: metadata constraints and ``callable(uuid)`` are
        evaluated by every part over its own rows."""
        if isinstance(filt, PartwiseFilter):
            filt = filt.spec
        if filt is None:
            return [None] * len(holders)
        if isinstance(filt, Mapping) or callable(filt):
            return [p._filter_mask(ns, filt) for _, p, ns in holders]
        raise ValueError("MultiGpuIndex filters are metadata constraints or callable(uuid) -> bool (row masks are per part)")

    # ------------------------------------------------------------------ the device step
    def _device_fanout(self, holders, queries: np.ndarray, k: int, filters):
        """All parts search at once, candidates meet on the first holder's device, ``mlv_merge_topk`` orders them.
        -> (dists f32 [nq,k], global rows i64 [nq,k], counts i32 [nq]) on the host."""
        import torch
        from .sharded import merge_topk_device
        nq = queries.shape[0]
        host_q = torch.from_numpy(np.ascontiguousarray(queries, dtype=np.float32))
        outs = []
        for (i, part, ns), flt in zip(holders, filters):
            dev = torch.device("cuda", part._device)
            with torch.cuda.device(dev):
                q = host_q.to(dev, non_blocking=True)
                d = torch.empty((nq, k), dtype=torch.float32, device=dev)
                r = torch.empty((nq, k), dtype=torch.int64, device=dev)
                c = torch.empty((nq,), dtype=torch.int32, device=dev)
                stream = torch.cuda.current_stream(dev).cuda_stream
                mask_ptr = 0
                prepared = flt if hasattr(flt, "_f") else None
                if flt is not None and prepared is None:       # a host mask: packed bitmap on the part's device
                    from .shard import pack_bitmap
                    words = torch.from_numpy(pack_bitmap(np.asarray(flt, dtype=bool)).view(np.int32)).to(dev)
                    mask_ptr = words.data_ptr()
                    outs.append(words)                          # keep it alive until the search has run
                with _bound(ns.shard, prepared):
                    ns.shard.search_device(q.data_ptr(), nq, k, d.data_ptr(), r.data_ptr(), c.data_ptr(), filter_ptr=mask_ptr,
                                           stream=stream)
                r = torch.where(r >= 0, r + (i << PART_SHIFT), r)
                outs.append((q, d, r))
        home = torch.device("cuda", holders[0][1]._device)
        triples = [o for o in outs if isinstance(o, tuple)]
        with torch.cuda.device(home):
            gd = torch.stack([d.to(home, non_blocking=True) for _, d, _ in triples])     # [G, nq, k]: peer copies
            gr = torch.stack([r.to(home, non_blocking=True) for _, _, r in triples])
            if len(triples) == 1:
                md, mr = gd[0], gr[0]
                mc = (mr >= 0).sum(dim=1).to(torch.int32)
            else:
                md, mr, mc = merge_topk_device(gd.contiguous(), gr.contiguous(), k)
            return md.cpu().numpy(), mr.cpu().numpy(), mc.cpu().numpy()

    # ------------------------------------------------------------------ IndexProtocol
    def _check_dim(self, namespace: str, dim: int) -> None:
        """The namespace's dimension is fixed by its first block on ANY part: a part that does not hold the namespace yet
        would otherwise create it with whatever dimension its slice has (hnswlib's add_items error, reference
        index.py:65 / GpuIndex.add)."""
        have = self.dimension(namespace)
        if have is not None and int(dim) != have:
            raise RuntimeError("Wrong dimensionality of the vectors")

    def add(self, vectors: Iterable[VectorProtocol], namespace: str) -> None:
        """reference index.py:50-67; the block is split over the parts (fewest live rows first)."""
        vectors = list(vectors)
        if not vectors:
            return
        dims = {int(np.asarray(v.values).shape[0]) for v in vectors}
        if len(dims) != 1:
            raise RuntimeError("Wrong dimensionality of the vectors")
        self._check_dim(namespace, dims.pop())
        sizes = split_block(len(vectors), [self._live(p, namespace) for p in self._parts])
        at = 0
        for part, n in zip(self._parts, sizes):
            if n:
                part.add(vectors[at:at + n], namespace)
                at += n

    def remove(self, ids: Sequence[UUID], namespace: str) -> None:
        """reference index.py:69-89: every part drops the ids it holds (unknown ids are ignored there too)."""
        for part in self._parts:
            part.remove(ids, namespace)

    def search(self, query: VectorDTO, top_k: int, namespace: str, metric: str, filter=None) -> List[SearchResult]:  # noqa: A002
        """reference index.py:91-129 (``filter``: metadata constraints or ``callable(uuid)``, additive)."""
        holders = [h for h in self._holders(namespace) if self._live(h[1], namespace) > 0]
        active = sum(self._live(p, namespace) for _, p, _ in holders)
        top_k = min(int(top_k), active)
        if top_k < 1:
            return []
        q = np.asarray(query.values, dtype=np.float32).reshape(-1)
        if any(q.shape[0] != ns.dim for _, _, ns in holders):
            return []                                  # the reference swallows the dimension error into [] (index.py:110-119)
        dists, rows, counts = self._fanout(holders, q[None, :], top_k, self._part_filters(holders, filter))
        return self._results(namespace, dists[0], rows[0], int(counts[0]), metric)

    def rebuild(self, source: Mapping[str, Iterable[VectorProtocol]], metric: str) -> None:
        """reference index.py:131-162: everything is replaced by ``source``, spread evenly over the parts."""
        per_part: List[Dict[str, list]] = [{} for _ in self._parts]
        for namespace, vectors in source.items():
            vectors = list(vectors)
            at = 0
            for part_src, n in zip(per_part, split_block(len(vectors), [0] * len(self._parts))):
                if n:
                    part_src[namespace] = vectors[at:at + n]
                    at += n
        for part, src in zip(self._parts, per_part):
            part.rebuild(src, metric)

    def is_rebuild_required(self, namespace: str) -> bool:
        """reference index.py:164-165"""
        return any(p.is_rebuild_required(namespace) for p in self._parts)

    # ------------------------------------------------------------------ additive surface (as GpuIndex)
    def dimension(self, namespace: str) -> Optional[int]:
        for _, _, ns in self._holders(namespace):
            return ns.dim
        return None

    def add_matrix(self, matrix: np.ndarray, namespace: str, ids: Optional[Sequence[UUID]] = None,
                   columns: Optional[Mapping[str, Sequence]] = None,
                   metadata: Optional[Sequence[Optional[Mapping]]] = None) -> np.ndarray:
        """Bulk ingest: the matrix is cut into one block per part; ``columns`` / ``metadata`` (as
        ``GpuIndex.add_matrix``) are cut the same way.  Returns the rows' UUID bytes [n, 16]."""
        data = np.ascontiguousarray(matrix, dtype=np.float32)
        if data.ndim != 2:
            raise ValueError("matrix must be [n, dim]")
        if data.shape[0] == 0:
            return np.empty((0, 16), dtype=np.uint8)
        self._check_dim(namespace, data.shape[1])
        if ids is not None and len(ids) != data.shape[0]:
            raise ValueError("len(ids) != rows")
        if metadata is not None and len(metadata) != data.shape[0]:
            raise ValueError("len(metadata) != rows")
        for name, values in (columns or {}).items():
            if len(values) != data.shape[0]:
                raise ValueError(f"column {name!r} has {len(values)} values for {data.shape[0]} rows")
        out = np.empty((data.shape[0], 16), dtype=np.uint8)
        at = 0
        for part, n in zip(self._parts, split_block(data.shape[0], [self._live(p, namespace) for p in self._parts])):
            if n:
                cols = {key: np.asarray(v)[at:at + n] for key, v in (columns or {}).items()}
                out[at:at + n] = part.add_matrix(data[at:at + n], namespace, ids=ids[at:at + n] if ids is not None else None,
                                                 columns=cols or None,
                                                 metadata=metadata[at:at + n] if metadata is not None else None)
                at += n
        return out

    def search_batch(self, queries: np.ndarray, top_k: int, namespace: str, metric: Optional[str] = None, filter=None):  # noqa: A002
        """-> (global rows i64 [nq,k] (-1 padded; ``uuids_of`` decodes them), scores f32 [nq,k], counts i32 [nq])."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        nq = q.shape[0]
        holders = [h for h in self._holders(namespace) if self._live(h[1], namespace) > 0]
        k = min(int(top_k), sum(self._live(p, namespace) for _, p, _ in holders))
        if k < 1 or any(q.shape[1] != ns.dim for _, _, ns in holders):
            return (np.full((nq, 0), -1, np.int64), np.empty((nq, 0), np.float32), np.zeros(nq, np.int32))
        dists, rows, counts = self._fanout(holders, q, k, self._part_filters(holders, filter))
        if (metric if metric is not None else self._space) == "cosine":
            dists = (1.0 - dists.astype(np.float64)).astype(np.float32)
        return rows, dists, counts

    def uuids_of(self, namespace: str, rows: np.ndarray) -> List[Optional[UUID]]:
        out = []
        for r in np.asarray(rows).reshape(-1).tolist():
            out.append(self._parts[r >> PART_SHIFT]._ns[namespace].uuid_of(r & PART_MASK) if r >= 0 else None)
        return out

    def range_search(self, query: VectorDTO, radius: float, namespace: str, metric: str, filter=None) -> List[SearchResult]:  # noqa: A002
        """Every live row of every part with hnswlib-form distance <= radius, nearest first.  Each part's hit list is
        ordered on its device; the lists are concatenated and ordered by ``mlv_index_order_pairs_device``."""
        holders = [h for h in self._holders(namespace) if self._live(h[1], namespace) > 0]
        if not holders:
            return []
        q = np.asarray(query.values, dtype=np.float32).reshape(-1)
        if any(q.shape[0] != ns.dim for _, _, ns in holders):
            return []
        lists = []
        for (i, part, ns), flt in zip(holders, self._part_filters(holders, filter)):
            (d, r), = ns.shard.range_search(q[None, :], float(radius), flt)
            lists.append((d, r + (i << PART_SHIFT)))
        dists = np.concatenate([d for d, _ in lists])
        rows = np.concatenate([r for _, r in lists])
        if len(lists) > 1 and len(rows):
            dists, rows = self._order_pairs(holders[0], dists, rows)
        return self._results(namespace, dists, rows, len(rows), metric)

    def _device_order_pairs(self, holder, dists: np.ndarray, rows: np.ndarray):
        """(distance, global row) pairs -> ascending, on the holder's device.  The part number rides in the row's high
        bits, which the 32-bit key cannot carry: pairs are keyed by their position in the concatenation (parts in
        order, each list ascending), which orders ties exactly like (part, row)."""
        import torch
        _, part, ns = holder
        dev = torch.device("cuda", part._device)
        with torch.cuda.device(dev):
            d = torch.from_numpy(np.ascontiguousarray(dists, np.float32)).to(dev)
            pos = torch.arange(len(rows), dtype=torch.int64, device=dev)
            od, op = torch.empty_like(d), torch.empty_like(pos)
            ns.shard.order_pairs_device(d.data_ptr(), pos.data_ptr(), len(rows), od.data_ptr(), op.data_ptr(),
                                        stream=torch.cuda.current_stream(dev).cuda_stream)
            order = op.cpu().numpy()
            return od.cpu().numpy(), rows[order]

    def where(self, namespace: str, constraints: Mapping) -> Optional[PartwiseFilter]:
        """Metadata constraints every holding part can decide on its device columns, or None (``GpuQueryProcessor`` then
        evaluates the predicate on the host and calls ``prepare_filter``)."""
        holders = self._holders(namespace)
        if not holders or any(p.where(namespace, constraints) is None for _, p, _ in holders):
            return None
        return PartwiseFilter(dict(constraints))

    def prepare_filter(self, namespace: str, filter) -> PartwiseFilter:  # noqa: A002
        if not (isinstance(filter, (Mapping, PartwiseFilter)) or callable(filter)):
            raise ValueError("MultiGpuIndex filters are metadata constraints or callable(uuid) -> bool (row masks are per part)")
        return filter if isinstance(filter, PartwiseFilter) else PartwiseFilter(filter)

    def info(self, namespace: str) -> dict:
        parts = [p.info(namespace) for _, p, _ in self._holders(namespace)]
        if not parts:
            raise KeyError(namespace)
        out = {key: sum(p[key] for p in parts) for key in ("rows", "live", "capacity", "device_bytes", "tombstones")}
        out.update(dim=parts[0]["dim"], space=parts[0]["space"], devices=[p["device"] for p in parts],
                   rows_per_device=[p["rows"] for p in parts])
        return out

    def metadata_columns(self, namespace: str) -> List[str]:
        names: List[str] = []
        for _, p, _ in self._holders(namespace):
            names += [n for n in p.metadata_columns(namespace) if n not in names]
        return names

    def namespaces(self) -> List[str]:
        seen: List[str] = []
        for p in self._parts:
            seen += [n for n in p.namespaces() if n not in seen]
        return seen

    def close(self) -> None:
        for p in self._parts:
            p.close()
