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
