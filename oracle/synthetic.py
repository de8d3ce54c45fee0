"""Deterministic synthetic rows / queries / filter buckets.  TEST ORACLE side of the generator.

The same counter-based generator exists three times and must agree bit for bit:
``oracle/exact_scan.c::orc_fill_synthetic`` (C), this file (numpy) and the CUDA kernel behind
``mlv_index_add_synthetic`` (product, ``mlvectordb_b200/csrc``).  It replaces SURVEY.md section
8d's ``default_rng`` proposal because a 10M x 768 matrix must be produced on the device in
seconds and regenerated chunk-wise by a streaming oracle (hard part H7).

    splitmix64(x): z = x + 0x9E3779B97F4A7C15; z = (z ^ z>>30) * 0xBF58476D1CE4E5B9;
                   z = (z ^ z>>27) * 0x94D049BB133111EB; return z ^ z>>31
    elem(row, col)  = float32(splitmix64(K + row*d + col) >> 40) * 2^-23 - 1      in [-1, 1)
    scale(row)      = 0.5 + float32(splitmix64(K2 + row) >> 40) * 2^-24            in [0.5, 1.5)
    K = splitmix64(seed), K2 = splitmix64(seed ^ 0xA5A5A5A5A5A5A5A5)
    value = elem * scale (one fp32 multiply) when ``scaled`` else elem
"""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = (np.asarray(x, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _key(seed: int) -> np.uint64:
    return splitmix64(np.array([seed & 0xFFFFFFFFFFFFFFFF], dtype=np.uint64))[0]


def rows(seed: int, first_row: int, n: int, d: int, scaled: bool = False) -> np.ndarray:
    """numpy restatement (slow; use ``oracle.cscan.fill_synthetic`` for big chunks)."""
    with np.errstate(over="ignore"):
        key = _key(seed)
        r = np.arange(first_row, first_row + n, dtype=np.uint64)[:, None]
        c = np.arange(d, dtype=np.uint64)[None, :]
        h = splitmix64(key + r * np.uint64(d) + c)
        m = (h >> np.uint64(40)).astype(np.uint32).astype(np.float32)
        v = m * np.float32(2.0 ** -23) - np.float32(1.0)
        if scaled:
            key2 = _key(seed ^ 0xA5A5A5A5A5A5A5A5)
            hs = splitmix64(key2 + r[:, 0])
            s = np.float32(0.5) + (hs >> np.uint64(40)).astype(np.uint32).astype(np.float32) * np.float32(2.0 ** -24)
            v = (v * s[:, None]).astype(np.float32)
        return v.astype(np.float32)


QUERY_SEED_OFFSET = 1_000_003


def queries(seed: int, nq: int, d: int) -> np.ndarray:
    """Queries come from an independent stream of the same generator."""
    return rows(seed + QUERY_SEED_OFFSET, 0, nq, d, scaled=False)


def buckets(seed: int, first_row: int, n: int) -> np.ndarray:
    """Filter column for config 4: bucket_i in [0, 100); predicate ``bucket < 100*s``."""
    with np.errstate(over="ignore"):
        key = _key(seed ^ 0x5EED5EED5EED5EED)
        r = np.arange(first_row, first_row + n, dtype=np.uint64)
        return (splitmix64(key + r) % np.uint64(100)).astype(np.int32)


def bitmap_from_mask(mask: np.ndarray) -> np.ndarray:
    """Pack a boolean row mask into the uint32 LSB-first bitmap the C-ABI takes."""
    mask = np.asarray(mask, dtype=bool)
    n = mask.shape[0]
    words = (n + 31) // 32
    padded = np.zeros(words * 32, dtype=bool)
    padded[:n] = mask
    return np.packbits(padded.reshape(-1, 8), axis=1, bitorder="little").reshape(-1).view(np.uint32).copy()
