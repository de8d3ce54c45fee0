"""Exact (brute-force) restatement of hnswlib 0.8.0's distance arithmetic.  TEST ORACLE.

PARITY UNPINNED (see ``oracle/__init__.py``): hnswlib is absent from this image and the
reference holds no golden vectors, so this file restates the *published* algorithm:

* ``l2``      d(x, q) = sum_i (x_i - q_i)^2          (squared; hnswlib ``space_l2.h`` L2Sqr)
* ``ip``      d(x, q) = 1 - sum_i x_i q_i             (hnswlib ``space_ip.h`` InnerProductDistance)
* ``cosine``  rows and queries are normalised in fp32 with ``1 / (sqrt(sum v_i^2) + 1e-30)``
              when they enter the index / the query (hnswlib ``bindings.cpp`` normalize_vector),
              then the ``ip`` distance is used.

These are the definitions the reference selects with ``hnswlib.Index(space=metric, dim)`` at
reference ``src/mlvectordb/implementations/index.py:36`` and consumes at ``index.py:111-128``.
All arithmetic is fp32 like hnswlib's; an optional fp64 shadow score adjudicates ties.

Ordering contract (reference ``index.py:121-128`` iterates hnswlib's ascending-distance
output): results are ascending by ``(distance, label)``.  hnswlib's own tie order is
unspecified (graph traversal); the label tie-break is this build's definition.

The additive features (filter bitmap, range search) have no reference code at all
(README-only, SURVEY.md section 8a last row); their semantics are *defined* here:

* filter: hnswlib-0.8 ``knn_query(..., filter=callable(label)->bool)`` semantics -- only
  labels that pass are candidates, k is clamped to the number of live-and-passing rows;
* range:  every live (and passing) row with hnswlib-form distance ``d <= radius``,
  ascending ``(d, label)``.
"""
from __future__ import annotations

from typing import Iterable, Iterator, Optional, Sequence, Tuple

import numpy as np

SPACES = ("l2", "ip", "cosine")
# aliases accepted additively by the product; the reference only knows hnswlib's.
ALIASES = {"euclidean": "l2", "dot": "ip", "inner_product": "ip", "cos": "cosine"}


def canonical_space(space: str) -> str:
    s = ALIASES.get(space, space)
    if s not in SPACES:
        raise ValueError(f"unknown space {space!r}")
    return s


def normalize_rows(x: np.ndarray) -> np.ndarray:
    """hnswlib bindings.cpp::normalize_vector, row-wise, fp32 throughout.

    norm = sum v_i^2 (fp32) ; inv = 1.0f / (sqrtf(norm) + 1e-30f) ; out_i = v_i * inv
    """
    x = np.ascontiguousarray(x, dtype=np.float32)
    if x.ndim == 1:
        x = x[None, :]
    sq = np.einsum("ij,ij->i", x, x, dtype=np.float32)
    inv = (np.float32(1.0) / (np.sqrt(sq, dtype=np.float32) + np.float32(1e-30))).astype(np.float32)
    return (x * inv[:, None]).astype(np.float32)


def distances(rows: np.ndarray, q: np.ndarray, space: str) -> np.ndarray:
    """fp32 hnswlib-form distances of one (already normalised, for cosine) query to rows."""
    rows = np.asarray(rows, dtype=np.float32)
    q = np.asarray(q, dtype=np.float32)
    if space == "l2":
        diff = rows - q[None, :]
        return np.einsum("ij,ij->i", diff, diff, dtype=np.float32)
    # ip and cosine (cosine data is normalised at add / query time)
    dot = rows @ q
    return (np.float32(1.0) - dot.astype(np.float32)).astype(np.float32)


def distances_f64(rows: np.ndarray, q: np.ndarray, space: str) -> np.ndarray:
    """fp64 shadow of :func:`distances` on the same fp32 inputs (tie / tolerance adjudication)."""
    r = np.asarray(rows, dtype=np.float64)
    qq = np.asarray(q, dtype=np.float64)
    if space == "l2":
        diff = r - qq[None, :]
        return np.einsum("ij,ij->i", diff, diff)
    return 1.0 - r @ qq


def _topk_merge(best_d: np.ndarray, best_l: np.ndarray, d: np.ndarray, l: np.ndarray, k: int):
    """Keep the k smallest by (distance, label) of the union of two candidate sets."""
    if best_d.size:
        d = np.concatenate([best_d, d])
        l = np.concatenate([best_l, l])
    if d.size > k:
        # partition on distance first (cheap), keep everything tied with the k-th
        kth = np.partition(d, k - 1)[k - 1]
        keep = d <= kth
        d, l = d[keep], l[keep]
    order = np.lexsort((l, d))[:k]
    return d[order], l[order]


def knn_stream(
    chunks: Iterable[Tuple[int, np.ndarray]],
    queries: np.ndarray,
    k: int,
    space: str,
    allow: Optional[np.ndarray] = None,
    prenormalized: bool = False,
) -> Tuple[list, list]:
    """Exact top-k over row chunks without ever holding a distance matrix.

    ``chunks`` yields ``(first_label, rows[n, d] fp32)``.  ``allow`` is an optional boolean
    mask over labels (tombstones AND filter already combined; False = not a candidate).
    Returns ``(labels_per_query, dists_per_query)``: python lists of 1-d arrays of length
    ``min(k, #candidates)``, ascending ``(distance, label)``.
    """
    space = canonical_space(space)
    q = np.ascontiguousarray(queries, dtype=np.float32)
    if q.ndim == 1:
        q = q[None, :]
    if space == "cosine":
        q = normalize_rows(q)
    nq = q.shape[0]
    best_d = [np.empty(0, np.float32) for _ in range(nq)]
    best_l = [np.empty(0, np.int64) for _ in range(nq)]
    for first, rows in chunks:
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        if space == "cosine" and not prenormalized:
            rows = normalize_rows(rows)
        labels = np.arange(first, first + rows.shape[0], dtype=np.int64)
        if allow is not None:
            m = np.asarray(allow[first:first + rows.shape[0]], dtype=bool)
            if not m.all():
                rows, labels = rows[m], labels[m]
        if rows.shape[0] == 0:
            continue
        for i in range(nq):
            d = distances(rows, q[i], space)
            best_d[i], best_l[i] = _topk_merge(best_d[i], best_l[i], d, labels, k)
    return best_l, best_d


def knn(rows: np.ndarray, queries: np.ndarray, k: int, space: str,
        allow: Optional[np.ndarray] = None, chunk_rows: int = 262144):
    """Exact top-k over an in-memory matrix (chunked so fp32 temporaries stay small)."""
    rows = np.asarray(rows)

    def gen() -> Iterator[Tuple[int, np.ndarray]]:
        for s in range(0, rows.shape[0], chunk_rows):
            yield s, rows[s:s + chunk_rows]

    return knn_stream(gen(), queries, k, space, allow=allow)


def range_stream(
    chunks: Iterable[Tuple[int, np.ndarray]],
    queries: np.ndarray,
    radius: float,
    space: str,
    allow: Optional[np.ndarray] = None,
    prenormalized: bool = False,
):
    """All candidates with hnswlib-form distance <= radius, ascending (distance, label)."""
    space = canonical_space(space)
    q = np.ascontiguousarray(queries, dtype=np.float32)
    if q.ndim == 1:
        q = q[None, :]
    if space == "cosine":
        q = normalize_rows(q)
    nq = q.shape[0]
    acc_d = [[] for _ in range(nq)]
    acc_l = [[] for _ in range(nq)]
    r32 = np.float32(radius)
    for first, rows in chunks:
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        if space == "cosine" and not prenormalized:
            rows = normalize_rows(rows)
        labels = np.arange(first, first + rows.shape[0], dtype=np.int64)
        if allow is not None:
            m = np.asarray(allow[first:first + rows.shape[0]], dtype=bool)
            if not m.all():
                rows, labels = rows[m], labels[m]
        if rows.shape[0] == 0:
            continue
        for i in range(nq):
            d = distances(rows, q[i], space)
            hit = d <= r32
            if hit.any():
                acc_d[i].append(d[hit])
                acc_l[i].append(labels[hit])
    out_l, out_d = [], []
    for i in range(nq):
        if acc_d[i]:
            d = np.concatenate(acc_d[i])
            l = np.concatenate(acc_l[i])
            order = np.lexsort((l, d))
            out_d.append(d[order])
            out_l.append(l[order])
        else:
            out_d.append(np.empty(0, np.float32))
            out_l.append(np.empty(0, np.int64))
    return out_l, out_d


def range_search(rows: np.ndarray, queries: np.ndarray, radius: float, space: str,
                 allow: Optional[np.ndarray] = None, chunk_rows: int = 262144):
    rows = np.asarray(rows)

    def gen():
        for s in range(0, rows.shape[0], chunk_rows):
            yield s, rows[s:s + chunk_rows]

    return range_stream(gen(), queries, radius, space, allow=allow)


# --------------------------------------------------------------------------------------
# Parity adjudication (SURVEY.md section 8d "Parity check", hard part H3)
# --------------------------------------------------------------------------------------
RTOL = 1e-5
ATOL = 1e-6


def scores_close(a, b, rtol: float = RTOL, atol: float = ATOL) -> np.ndarray:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) <= rtol * np.maximum(np.abs(a), np.abs(b)) + atol


def check_topk_parity(got_labels: Sequence[int], got_scores: Sequence[float],
                      ref_labels: Sequence[int], ref_scores: Sequence[float],
                      all_ref_scores=None, rtol: float = RTOL, atol: float = ATOL) -> Optional[str]:
    """Return None when ``got`` matches the oracle's top-k, else a message.

    Rule (BASELINE.json north_star): id sets equal except swaps among candidates whose
    oracle scores lie within tolerance of the k-th oracle score; scores within
    ``rtol*max(|a|,|b|)+atol`` position by position; output ascending.

    ``all_ref_scores``: optional callable ``label -> oracle score`` used to adjudicate an id
    that is in ``got`` but not in the oracle's list (it must tie with the k-th).
    """
    got_labels = [int(x) for x in got_labels]
    ref_labels = [int(x) for x in ref_labels]
    got_scores = np.asarray(got_scores, dtype=np.float64)
    ref_scores = np.asarray(ref_scores, dtype=np.float64)
    if len(got_labels) != len(ref_labels):
        return f"count {len(got_labels)} != oracle {len(ref_labels)}"
    if len(got_labels) == 0:
        return None
    if len(set(got_labels)) != len(got_labels):
        return "duplicate labels in result"
    ok = scores_close(got_scores, ref_scores, rtol, atol)
    if not ok.all():
        j = int(np.argmin(ok))
        return f"score[{j}] {got_scores[j]!r} vs oracle {ref_scores[j]!r}"
    tol_sorted = rtol * np.abs(got_scores[:-1]) + atol
    if (np.diff(got_scores) < -tol_sorted).any():
        return "scores not ascending"
    kth = ref_scores[-1]
    tol = rtol * abs(kth) + atol
    ref_pos = {l: i for i, l in enumerate(ref_labels)}
    for l, s in zip(got_labels, got_scores):
        if l in ref_pos:
            continue
        # not in the oracle's list: admissible only as a tie with the k-th oracle score
        if all_ref_scores is not None:
            s_ref = float(all_ref_scores(l))
            if abs(s_ref - kth) > 2 * tol:
                return f"label {l} (oracle score {s_ref}) is not within tolerance of k-th {kth}"
        elif abs(s - kth) > 2 * tol:
            return f"label {l} (score {s}) not in oracle set and not tied with k-th {kth}"
    got_set = set(got_labels)
    for l, s in zip(ref_labels, ref_scores):
        if l not in got_set and abs(s - kth) > 2 * tol:
            return f"oracle label {l} (score {s}) missing and not tied with k-th {kth}"
    return None


# --------------------------------------------------------------------------------------
# Columnar metadata predicates (no reference code: README.md:123,477 and the stale client's
# ``filter`` dict of equality constraints, examples/api_client.py:65-74).  Semantics of
# ``mlv_filter_create_where`` (include/mlv_index.h) are *defined* here.
# --------------------------------------------------------------------------------------
COLUMN_MISSING = -(2 ** 31)

_WHERE_OPS = {
    "==": lambda v, a, b: v == a,
    "!=": lambda v, a, b: v != a,
    "<": lambda v, a, b: v < a,
    "<=": lambda v, a, b: v <= a,
    ">": lambda v, a, b: v > a,
    ">=": lambda v, a, b: v >= a,
    "between": lambda v, a, b: (v >= a) & (v <= b),
}


def where_mask(columns, predicates, n_rows: int) -> np.ndarray:
    """bool[n_rows]: row passes iff for EVERY predicate ``(column, op, a[, b])`` its int32 value is
    not ``COLUMN_MISSING`` and ``value op a`` holds.  ``columns``: mapping column -> int32 array
    (shorter than ``n_rows`` or absent = missing for the remaining rows)."""
    mask = np.ones(n_rows, dtype=bool)
    for p in predicates:
        column, op, a = p[0], p[1], int(p[2])
        b = int(p[3]) if len(p) > 3 else 0
        v = np.full(n_rows, COLUMN_MISSING, dtype=np.int64)
        col = columns.get(column)
        if col is not None:
            col = np.asarray(col, dtype=np.int64)[:n_rows]
            v[: col.shape[0]] = col
        mask &= (v != COLUMN_MISSING) & _WHERE_OPS[op](v, a, b)
    return mask
