"""ctypes loader for the plain-C oracle (``oracle/exact_scan.c``).  TEST ORACLE / CPU BASELINE.

PARITY UNPINNED (``oracle/__init__.py``).  Builds ``oracle/_build/liborcscan.so`` with the
committed ``oracle/Makefile`` when it is missing or older than the source.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liborcscan.so")
_SRC = os.path.join(_HERE, "exact_scan.c")
_lib = None

SPACE_CODE = {"l2": 0, "ip": 1, "cosine": 2}


def build(force: bool = False) -> str:
    stale = (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(_SRC)
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "_build/liborcscan.so"])
    return _SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        f32p, i64p, i32p, u32p = (C.POINTER(C.c_float), C.POINTER(C.c_int64),
                                  C.POINTER(C.c_int32), C.POINTER(C.c_uint32))
        L.orc_l2sqr.restype = C.c_float
        L.orc_l2sqr.argtypes = [f32p, f32p, C.c_size_t]
        L.orc_ip.restype = C.c_float
        L.orc_ip.argtypes = [f32p, f32p, C.c_size_t]
        L.orc_l2sqr_simd16.restype = C.c_float
        L.orc_l2sqr_simd16.argtypes = [f32p, f32p, C.c_size_t]
        L.orc_ip_simd16.restype = C.c_float
        L.orc_ip_simd16.argtypes = [f32p, f32p, C.c_size_t]
        L.orc_normalize.restype = None
        L.orc_normalize.argtypes = [f32p, f32p, C.c_size_t, C.c_size_t]
        L.orc_distances.restype = None
        L.orc_distances.argtypes = [f32p, C.c_size_t, C.c_size_t, f32p, C.c_int, C.c_int, f32p]
        L.orc_knn.restype = None
        L.orc_knn.argtypes = [f32p, C.c_size_t, C.c_size_t, f32p, C.c_size_t, C.c_int, C.c_int, C.c_int,
                              u32p, C.c_int64, i64p, f32p, i32p]
        L.orc_fill_synthetic.restype = None
        L.orc_fill_synthetic.argtypes = [f32p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_threads.restype = None
        L.orc_set_threads.argtypes = [C.c_int]
        u64p = C.POINTER(C.c_uint64)
        L.orc_knn_synthetic.restype = None
        L.orc_knn_synthetic.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, f32p, C.c_size_t, C.c_int,
                                        C.c_int, C.c_int, u32p, i64p, f32p, i32p]
        L.orc_range_synthetic.restype = None
        L.orc_range_synthetic.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, f32p, C.c_size_t, C.c_float,
                                          C.c_int, C.c_int, u32p, C.c_uint64, i64p, f32p, u64p]
        L.orc_distances_synthetic.restype = None
        L.orc_distances_synthetic.argtypes = [C.c_uint64, i64p, C.c_size_t, C.c_uint64, C.c_int, f32p, C.c_int, C.c_int, f32p]
        _lib = L
    return _lib


def _f32(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def normalize(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    lib().orc_normalize(_f32(x), _f32(out), x.shape[0], x.shape[1])
    return out


def distances(rows: np.ndarray, q: np.ndarray, space: str, simd16: bool = False) -> np.ndarray:
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    q = np.ascontiguousarray(q, dtype=np.float32)
    out = np.empty(rows.shape[0], np.float32)
    lib().orc_distances(_f32(rows), rows.shape[0], rows.shape[1], _f32(q), SPACE_CODE[space], int(simd16), _f32(out))
    return out


def knn(rows: np.ndarray, queries: np.ndarray, k: int, space: str, allow_bitmap: np.ndarray | None = None,
        simd16: bool = True, first_label: int = 0):
    """Rows / queries must already be normalised for cosine (use :func:`normalize`)."""
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    queries = np.ascontiguousarray(queries, dtype=np.float32)
    if queries.ndim == 1:
        queries = queries[None, :]
    nq = queries.shape[0]
    labels = np.empty((nq, k), np.int64)
    dists = np.empty((nq, k), np.float32)
    counts = np.empty(nq, np.int32)
    bm = None
    if allow_bitmap is not None:
        allow_bitmap = np.ascontiguousarray(allow_bitmap, dtype=np.uint32)
        bm = allow_bitmap.ctypes.data_as(C.POINTER(C.c_uint32))
    lib().orc_knn(_f32(rows), rows.shape[0], rows.shape[1], _f32(queries), nq, k, SPACE_CODE[space], int(simd16),
                  bm, first_label, labels.ctypes.data_as(C.POINTER(C.c_int64)), _f32(dists),
                  counts.ctypes.data_as(C.POINTER(C.c_int32)))
    return labels, dists, counts


def fill_synthetic(seed: int, first_row: int, n: int, d: int, scaled: bool) -> np.ndarray:
    out = np.empty((n, d), np.float32)
    lib().orc_fill_synthetic(_f32(out), seed, first_row, n, d, int(scaled))
    return out


def num_threads() -> int:
    return int(lib().orc_num_threads())


def host_threads() -> int:
    """Cores this process may run on (cgroup / affinity aware)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def use_all_cores() -> int:
    """Run the OpenMP loops on every core the process may use, whatever OMP_NUM_THREADS says
    (``torch.distributed.run`` exports OMP_NUM_THREADS=1 to its workers).  Returns the thread count."""
    n = host_threads()
    lib().orc_set_threads(n)
    return num_threads()


def _prep_queries(queries: np.ndarray, space: str) -> np.ndarray:
    q = np.ascontiguousarray(queries, dtype=np.float32)
    if q.ndim == 1:
        q = q[None, :]
    return normalize(q) if space == "cosine" else q


def _bitmap(mask_or_words, n: int):
    if mask_or_words is None:
        return None, None
    a = np.asarray(mask_or_words)
    if a.dtype != np.uint32:
        m = np.zeros(((n + 31) // 32) * 32, dtype=bool)
        m[:n] = a.astype(bool)[:n]
        a = np.packbits(m, bitorder="little").view(np.uint32)
    a = np.ascontiguousarray(a)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint32))


def knn_synthetic(seed: int, first_row: int, n: int, d: int, scaled: bool, queries: np.ndarray, k: int, space: str,
                  allow=None, simd16: bool = True):
    """Streamed exact kNN over generator rows ``first_row .. first_row+n-1`` (never materialised).
    ``queries`` are RAW (normalised here for cosine).  ``allow``: bool mask or uint32 bitmap over
    ``row - first_row``.  -> (labels i64 [nq,k] = generator rows, dists f32 [nq,k], counts i32 [nq])."""
    q = _prep_queries(queries, space)
    nq = q.shape[0]
    labels = np.empty((nq, k), np.int64)
    dists = np.empty((nq, k), np.float32)
    counts = np.empty(nq, np.int32)
    keep, bm = _bitmap(allow, n)
    lib().orc_knn_synthetic(seed, first_row, n, d, int(scaled), _f32(q), nq, k, SPACE_CODE[space], int(simd16), bm,
                            labels.ctypes.data_as(C.POINTER(C.c_int64)), _f32(dists),
                            counts.ctypes.data_as(C.POINTER(C.c_int32)))
    del keep
    return labels, dists, counts


def range_synthetic(seed: int, first_row: int, n: int, d: int, scaled: bool, queries: np.ndarray, radius: float,
                    space: str, allow=None, simd16: bool = True, max_hits: int = 4096):
    """Streamed range search over generator rows: per query (labels, dists) ascending (distance, label)."""
    q = _prep_queries(queries, space)
    nq = q.shape[0]
    keep, bm = _bitmap(allow, n)
    while True:
        labels = np.empty((nq, max_hits), np.int64)
        dists = np.empty((nq, max_hits), np.float32)
        counts = np.zeros(nq, np.uint64)
        lib().orc_range_synthetic(seed, first_row, n, d, int(scaled), _f32(q), nq, C.c_float(radius), SPACE_CODE[space],
                                  int(simd16), bm, max_hits, labels.ctypes.data_as(C.POINTER(C.c_int64)), _f32(dists),
                                  counts.ctypes.data_as(C.POINTER(C.c_uint64)))
        if int(counts.max()) <= max_hits:
            break
        max_hits = int(counts.max())
    del keep
    return [(labels[i, :int(counts[i])].copy(), dists[i, :int(counts[i])].copy()) for i in range(nq)]


def distances_synthetic(seed: int, labels, d: int, scaled: bool, query: np.ndarray, space: str, simd16: bool = True) -> np.ndarray:
    """Oracle distances of generator rows ``labels`` to ONE raw query."""
    q = _prep_queries(query, space)
    lab = np.ascontiguousarray(labels, dtype=np.int64)
    out = np.empty(lab.shape[0], np.float32)
    lib().orc_distances_synthetic(seed, lab.ctypes.data_as(C.POINTER(C.c_int64)), lab.shape[0], d, int(scaled), _f32(q[0]),
                                  SPACE_CODE[space], int(simd16), _f32(out))
    return out


def check_knn_synthetic(got_rows, got_dists, got_counts, seed: int, first_row: int, n: int, d: int, scaled: bool,
                        queries: np.ndarray, k: int, space: str, allow=None):
    """Full parity rule (``oracle.exact.check_topk_parity``) of a search result against the streamed oracle.
    Returns None or the first mismatch message."""
    from . import exact
    L, D, Cn = knn_synthetic(seed, first_row, n, d, scaled, queries, k, space, allow=allow)
    q = np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, d)
    for i in range(q.shape[0]):
        c = int(Cn[i])
        if int(got_counts[i]) != c:
            return f"query {i}: count {int(got_counts[i])} vs oracle {c}"
        msg = exact.check_topk_parity(
            got_rows[i, :c], got_dists[i, :c], L[i, :c], D[i, :c],
            all_ref_scores=lambda l, i=i: float(distances_synthetic(seed, [l], d, scaled, q[i], space)[0]))
        if msg is not None:
            return f"query {i}: {msg}"
    return None
