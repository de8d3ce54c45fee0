"""TEST-ONLY stand-in for ``mlvectordb_b200.shard.DeviceShard`` backed by the CPU oracle.

The CPU suite (``-m "not gpu"``) has no device, but most of ``GpuIndex`` / ``GpuQueryProcessor`` / ``GpuRestAPI`` /
``snapshot`` is host logic: id maps, counters, thresholds, compaction bookkeeping, metadata codecs, filter caching,
response shapes.  ``tests/test_host_logic_cpu.py`` monkeypatches this class in for the duration of a test so that
logic runs here; the arithmetic behind it is ``oracle/exact.py``.  Nothing under ``mlvectordb_b200/`` knows about
this file (``tests/test_capi_cpu.py::test_product_never_imports_the_oracle``)."""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np

from mlvectordb_b200.shard import PreparedFilter, canonical_space
from oracle import exact


class FakePrepared(PreparedFilter):
    def __init__(self, shard, mask):  # noqa: D401 -- no C handle behind it
        self._shard, self._mask, self._f = shard, np.asarray(mask, bool).copy(), SimpleNamespace(value=1)

    @property
    def passing(self) -> int:
        n = self._shard.rows
        m = np.zeros(n, bool)
        m[: min(n, self._mask.shape[0])] = self._mask[:n]
        return int((m & self._shard._live[:n]).sum())

    def bitmap(self, n_words=None):
        n = self._shard.rows
        m = np.zeros(((n + 31) // 32) * 32, bool)
        m[: min(n, self._mask.shape[0])] = self._mask[:n]
        return np.packbits(m, bitorder="little").view(np.uint32)

    def close(self) -> None:
        self._f = SimpleNamespace(value=0)

    def __del__(self):
        pass


class FakeShard:
    def __init__(self, dim, space="l2", capacity=0, device=0, row_base=0):
        self.space, self.dim, self.device, self.row_base = canonical_space(space), int(dim), int(device), int(row_base)
        self._x = np.empty((0, self.dim), np.float32)
        self._live = np.empty(0, bool)
        self._cols = {}
        self._h = SimpleNamespace(value=1)

    # -- lifecycle / info
    def close(self):
        self._h = SimpleNamespace(value=0)

    @property
    def rows(self):
        return self._x.shape[0]

    @property
    def live(self):
        return int(self._live.sum())

    def info(self):
        return SimpleNamespace(rows=self.rows, live=self.live, capacity=max(self.rows, 16), dim=self.dim, device=self.device,
                               device_bytes=self._x.nbytes, row_base=self.row_base)

    # -- mutation
    def add(self, rows):
        x = np.ascontiguousarray(rows, dtype=np.float32).reshape(-1, self.dim)
        first = self.rows
        self._x = np.concatenate([self._x, x])
        self._live = np.concatenate([self._live, np.ones(x.shape[0], bool)])
        return first

    def mark_deleted(self, rows):
        changed = 0
        for r in np.asarray(rows, dtype=np.int64):
            if 0 <= r < self.rows and self._live[r]:
                self._live[r] = False
                changed += 1
        return changed

    def compact(self):
        keep = self._live.copy()
        mapping = np.full(self.rows, -1, np.int64)
        mapping[keep] = np.arange(int(keep.sum()))
        self._x = self._x[keep]
        self._cols = {c: v[keep[: v.shape[0]]] if v.shape[0] == keep.shape[0] else v for c, v in self._cols.items()}
        self._live = np.ones(self._x.shape[0], bool)
        return mapping

    # -- filters
    def _mask_of(self, filt):
        if filt is None:
            return None
        if isinstance(filt, FakePrepared):
            m = np.zeros(self.rows, bool)
            m[: min(self.rows, filt._mask.shape[0])] = filt._mask[: self.rows]
            return m
        f = np.asarray(filt)
        if f.dtype == np.uint32:
            return np.unpackbits(f.view(np.uint8), bitorder="little")[: self.rows].astype(bool)
        if f.shape[0] != self.rows:
            raise ValueError(f"filter mask has {f.shape[0]} entries for {self.rows} rows")
        return f.astype(bool)

    def prepare_filter(self, filt):
        return FakePrepared(self, self._mask_of(filt))

    def set_column(self, column, values, first_row=0):
        v = np.asarray(values, dtype=np.int32)
        if first_row + v.shape[0] > self.rows:
            raise RuntimeError("column write beyond the stored rows")
        col = self._cols.get(column)
        if col is None or col.shape[0] < self.rows:
            grown = np.full(self.rows, exact.COLUMN_MISSING, np.int32)
            if col is not None:
                grown[: col.shape[0]] = col
            col = self._cols[column] = grown
        col[first_row:first_row + v.shape[0]] = v

    def get_column(self, column, first_row=0, n=None):
        n = self.rows - first_row if n is None else int(n)
        out = np.full(self.rows, exact.COLUMN_MISSING, np.int32)
        col = self._cols.get(column)
        if col is not None:
            out[: col.shape[0]] = col
        return out[first_row:first_row + n]

    def where(self, predicates):
        return FakePrepared(self, exact.where_mask(self._cols, predicates, self.rows))

    # -- queries
    def _allow(self, filt):
        m = self._mask_of(filt)
        return self._live if m is None else (self._live & m)

    def search(self, queries, k, filt=None):
        q = np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, self.dim)
        L, D = exact.knn(self._x, q, k, self.space, allow=self._allow(filt)) if self.rows else ([np.empty(0, np.int64)] * len(q), [np.empty(0, np.float32)] * len(q))
        d = np.full((len(q), k), np.inf, np.float32)
        r = np.full((len(q), k), -1, np.int64)
        c = np.zeros(len(q), np.int32)
        for i in range(len(q)):
            n = len(L[i])
            d[i, :n], r[i, :n], c[i] = D[i], np.asarray(L[i]) + self.row_base, n
        return d, r, c

    def submit(self, queries, k, exchange=False):
        return SimpleNamespace(result=lambda out=self.search(queries, k): out)

    def range_search(self, queries, radius, filt=None, max_hits=1024):
        q = np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, self.dim)
        L, D = exact.range_search(self._x, q, radius, self.space, allow=self._allow(filt))
        return [(np.asarray(D[i], np.float32), np.asarray(L[i], np.int64) + self.row_base) for i in range(len(q))]

    # -- snapshot (stored form == what add() received: this stand-in normalises at query time)
    def export_rows(self, first_row=0, n=None, out=None):
        n = self.rows - first_row if n is None else int(n)
        return self._x[first_row:first_row + n].copy()

    def export_live(self):
        m = np.zeros(((self.rows + 31) // 32) * 32, bool)
        m[: self.rows] = self._live
        return np.packbits(m, bitorder="little").view(np.uint32)

    def import_rows(self, rows, live_words=None):
        first = self.add(rows)
        if live_words is not None:
            bits = np.unpackbits(np.ascontiguousarray(live_words, dtype=np.uint32).view(np.uint8), bitorder="little")[: len(rows)].astype(bool)
            self._live[first:] = bits
        return first
