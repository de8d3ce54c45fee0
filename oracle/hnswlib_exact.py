"""Exact stand-in for the seven ``hnswlib.Index`` calls the reference makes.  TEST ORACLE.

PARITY UNPINNED (``oracle/__init__.py``).  Call sites this mirrors, all in reference
``src/mlvectordb/implementations/index.py``:

* ``hnswlib.Index(space=metric, dim=dim)``                           ``:36``
* ``init_index(max_elements=10_000, ef_construction=, M=)``          ``:37``
* ``set_ef(50)``                                                     ``:38``
* ``get_current_count()``                                            ``:56``
* ``add_items(float32[n, d], labels[n])``                            ``:65``, ``:158``
* ``mark_deleted(label)``                                            ``:80``
* ``knn_query(float32[1, d], k) -> (labels uint64, distances f32)``  ``:111``, ``:115``

Semantics follow hnswlib 0.8.0's python bindings: fp32 storage; ``cosine`` normalises at add
and at query; ``knn_query`` returns ascending distance, skips deleted labels, and raises
``RuntimeError`` when it cannot fill ``k`` rows or the dimension is wrong; ``add_items`` past
``max_elements`` raises ``RuntimeError``.  The search itself is *exact* (a full scan) -- the
graph/ANN part of hnswlib is out of scope.

``sys.modules["hnswlib"] = oracle.hnswlib_exact`` lets the unmodified reference wrappers run
(``oracle/refload.py``).  ``ENFORCE_MAX_ELEMENTS = False`` lifts the 10 000-row cap, which is
a reference limitation (``index.py:37``), not a semantic.
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np

from . import exact

ENFORCE_MAX_ELEMENTS = True


class Index:
    def __init__(self, space: str = "l2", dim: int = 0):
        if space not in exact.SPACES:
            raise RuntimeError(f"Space name must be one of l2, ip, or cosine. (got {space!r})")
        self.space = space
        self.dim = int(dim)
        self._normalize = space == "cosine"
        self._data = np.empty((0, self.dim), dtype=np.float32)
        self._labels = np.empty(0, dtype=np.int64)       # external label of each slot
        self._deleted = np.empty(0, dtype=bool)
        self._slot_of = {}                                # label -> slot
        self._count = 0
        self.max_elements = 0
        self.ef = 10
        self._initialised = False

    # -- construction ---------------------------------------------------------------
    def init_index(self, max_elements: int, ef_construction: int = 200, M: int = 16,
                   random_seed: int = 100, allow_replace_deleted: bool = False) -> None:
        if self._initialised:
            raise RuntimeError("The index is already initiated.")
        self.max_elements = int(max_elements)
        self.ef_construction = ef_construction
        self.M = M
        self._initialised = True

    def set_ef(self, ef: int) -> None:
        self.ef = int(ef)

    def get_current_count(self) -> int:
        return self._count

    def get_max_elements(self) -> int:
        return self.max_elements

    def resize_index(self, new_size: int) -> None:
        if new_size < self._count:
            raise RuntimeError("Cannot resize, max element is less than the current number of elements")
        self.max_elements = int(new_size)

    # -- mutation -------------------------------------------------------------------
    def _reserve(self, n: int) -> None:
        need = self._count + n
        cap = self._data.shape[0]
        if need <= cap:
            return
        new_cap = max(need, cap * 2, 16)
        data = np.empty((new_cap, self.dim), dtype=np.float32)
        data[: self._count] = self._data[: self._count]
        self._data = data
        lab = np.empty(new_cap, dtype=np.int64)
        lab[: self._count] = self._labels[: self._count]
        self._labels = lab
        dele = np.zeros(new_cap, dtype=bool)
        dele[: self._count] = self._deleted[: self._count]
        self._deleted = dele

    def add_items(self, data, ids=None, num_threads: int = -1, replace_deleted: bool = False) -> None:
        x = np.asarray(data, dtype=np.float32)
        if x.ndim == 1:
            x = x[None, :]
        if x.ndim != 2:
            raise RuntimeError("Input vector data wrong shape. Number of dimensions %d. Data must be a 1D or 2D array." % x.ndim)
        if x.shape[1] != self.dim:
            raise RuntimeError("Wrong dimensionality of the vectors")
        n = x.shape[0]
        if ids is None:
            ids = np.arange(self._count, self._count + n)
        ids = np.asarray(ids).reshape(-1)
        if ids.shape[0] != n:
            raise RuntimeError("Wrong dimensionality of the labels")
        if self._normalize:
            x = exact.normalize_rows(x)
        fresh = sum(1 for l in ids.tolist() if int(l) not in self._slot_of)
        if ENFORCE_MAX_ELEMENTS and self._count + fresh > self.max_elements:
            raise RuntimeError("The number of elements exceeds the specified limit")
        self._reserve(fresh)
        for row, l in zip(x, ids.tolist()):
            l = int(l)
            slot = self._slot_of.get(l)
            if slot is None:                       # hnswlib addPoint: new label -> new slot
                slot = self._count
                self._count += 1
                self._slot_of[l] = slot
                self._labels[slot] = l
                self._deleted[slot] = False
            elif self._deleted[slot]:              # hnswlib addPoint on a deleted label: undelete + update
                self._deleted[slot] = False
            self._data[slot] = row                 # existing label -> updatePoint

    def mark_deleted(self, label: int) -> None:
        slot = self._slot_of.get(int(label))
        if slot is None:
            raise RuntimeError("Label not found")
        if self._deleted[slot]:
            raise RuntimeError("The requested to delete element is already deleted")
        self._deleted[slot] = True

    def unmark_deleted(self, label: int) -> None:
        slot = self._slot_of.get(int(label))
        if slot is None:
            raise RuntimeError("Label not found")
        if not self._deleted[slot]:
            raise RuntimeError("The requested to undelete element is not deleted")
        self._deleted[slot] = False

    # -- query ----------------------------------------------------------------------
    def knn_query(self, data, k: int = 1, num_threads: int = -1,
                  filter: Optional[Callable[[int], bool]] = None):
        q = np.asarray(data, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise RuntimeError("Wrong dimensionality of the vectors")
        n = self._count
        allow = ~self._deleted[:n]
        if filter is not None:
            passing = np.fromiter((bool(filter(int(l))) for l in self._labels[:n]), dtype=bool, count=n)
            allow = allow & passing
        # slots play the label role inside exact.knn; map back afterwards
        slot_lists, dist_lists = exact.knn_stream(
            [(0, self._data[:n])], q, k, self.space, allow=allow, prenormalized=True)
        nq = q.shape[0]
        labels = np.empty((nq, k), dtype=np.uint64)
        dists = np.empty((nq, k), dtype=np.float32)
        for i in range(nq):
            if slot_lists[i].shape[0] < k:
                raise RuntimeError(
                    "Cannot return the results in a contigious 2D array. Probably ef or M is too small")
            ext = self._labels[slot_lists[i]]
            # ascending (distance, external label)
            order = np.lexsort((ext, dist_lists[i]))
            labels[i] = ext[order].astype(np.uint64)
            dists[i] = dist_lists[i][order]
        return labels, dists

    def get_items(self, ids, return_type: str = "numpy"):
        rows = np.stack([self._data[self._slot_of[int(l)]] for l in ids]) if len(ids) else np.empty((0, self.dim), np.float32)
        return rows if return_type == "numpy" else rows.tolist()

    def get_ids_list(self):
        return [int(l) for l in self._labels[: self._count]]
