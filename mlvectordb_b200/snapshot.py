"""Snapshots of a ``GpuIndex``.

The reference has no persistence: storage is in memory, a restart loses everything, and
``save_index`` / ``load_index`` / ``create_backup`` exist only as README prose (``README.md:240-241,276-277``);
its one recovery path is ``Index.rebuild`` from ``storage.namespace_map`` (``implementations/index.py:131-162``),
i.e. a full re-upload through per-row Python objects.  A snapshot writes every namespace's device matrix
*as stored* (cosine rows already normalised, tombstoned rows included, so row numbers -- the hnswlib
labels of ``index.py:56-63`` -- and the id map stay valid), the tombstone bitmap, the row -> UUID table,
the metadata columns and their codec; loading re-creates the matrices bit for bit with
``mlv_index_import_rows`` (no second normalisation), so searches before and after are identical.

Layout: a directory with ``manifest.json`` and, per namespace ``i`` of save generation ``g``, ``g<g>.ns<i>.rows.npy``
([rows, dim] fp32), ``g<g>.ns<i>.live.npy`` (uint32 bitmap, bit set = live), ``g<g>.ns<i>.ids.npy`` ([rows, 16] uint8),
``g<g>.ns<i>.col<j>.npy`` (int32).  Plain ``.npy`` files, streamed in 2^20-row chunks.  Saving over an existing snapshot
writes a NEW generation beside the old files and replaces the manifest atomically last, so a crash mid-save leaves the
previous snapshot loadable; the previous generation's files are removed after the manifest switch.
"""
from __future__ import annotations

import json
import os
from typing import Optional

import numpy as np

from .columns import ColumnCodec

FORMAT_VERSION = 1
CHUNK_ROWS = 1 << 20   # multiple of 32: chunks start on a bitmap word


def save_index(index, path: str) -> dict:
    """Write ``index`` (a ``GpuIndex``) under directory ``path``; returns the manifest."""
    os.makedirs(path, exist_ok=True)
    old_gen = None
    try:
        with open(os.path.join(path, "manifest.json")) as fh:
            old_gen = int(json.load(fh).get("generation", 0))
    except (OSError, ValueError):
        pass
    gen = 0 if old_gen is None else old_gen + 1
    for stale in os.listdir(path):          # leftovers of a save that crashed before its manifest switch
        if stale.startswith(f"g{gen}.") or stale == "manifest.json.tmp":
            os.remove(os.path.join(path, stale))
    manifest = {
        "generation": gen,
        "format": FORMAT_VERSION, "space": index._space, "rebuild_threshold": index._rebuild_threshold,
        "ef_construction": index._ef_construction, "M": index._M, "auto_compact": index._auto_compact,
        "namespaces": [],
    }
    for i, (name, ns) in enumerate(index._ns.items()):
        n = ns.n
        rows = np.lib.format.open_memmap(os.path.join(path, f"g{gen}.ns{i}.rows.npy"), mode="w+", dtype=np.float32, shape=(n, ns.dim))
        for r0 in range(0, n, CHUNK_ROWS):
            nr = min(CHUNK_ROWS, n - r0)
            rows[r0:r0 + nr] = ns.shard.export_rows(r0, nr)
        rows.flush()
        del rows
        np.save(os.path.join(path, f"g{gen}.ns{i}.live.npy"), ns.shard.export_live())
        np.save(os.path.join(path, f"g{gen}.ns{i}.ids.npy"), ns.ids[:n])
        codec = ns.codec.to_json()
        saved_cols = []
        for key, st in codec["columns"].items():
            if st["kind"] != "host":
                np.save(os.path.join(path, f"g{gen}.ns{i}.col{st['index']}.npy"), ns.shard.get_column(st["index"], 0, n))
                saved_cols.append(st["index"])
        manifest["namespaces"].append({
            "name": name, "dim": ns.dim, "space": ns.space, "rows": n, "total": ns.total, "deleted": ns.deleted,
            "rebuild_required": bool(ns.rebuild_required), "codec": codec, "columns": saved_cols,
        })
    tmp = os.path.join(path, "manifest.json.tmp")
    with open(tmp, "w") as fh:
        json.dump(manifest, fh)
    os.replace(tmp, os.path.join(path, "manifest.json"))   # the manifest switches last: a torn save leaves the old one valid
    if old_gen is not None:
        prefix = f"g{old_gen}." if "generation" in _peek_old(path, old_gen) else "ns"
        for stale in os.listdir(path):
            if stale.startswith(prefix) and stale.endswith(".npy") and not stale.startswith(f"g{gen}."):
                os.remove(os.path.join(path, stale))
    return manifest


def _peek_old(path: str, old_gen: int) -> dict:
    """{'generation': ..} when the previous snapshot used generation-prefixed file names (always, since format 1's
    second revision); manifests written before that have un-prefixed ``ns<i>.*`` files."""
    return {"generation": old_gen} if any(f.startswith(f"g{old_gen}.") for f in os.listdir(path)) else {}


def load_index(path: str, device: int = 0, index_cls=None):
    """Re-create a ``GpuIndex`` from ``save_index``'s directory on CUDA device ``device``."""
    if index_cls is None:
        from .index import GpuIndex as index_cls
    with open(os.path.join(path, "manifest.json")) as fh:
        manifest = json.load(fh)
    if manifest.get("format") != FORMAT_VERSION:
        raise ValueError(f"snapshot format {manifest.get('format')!r} != {FORMAT_VERSION}")
    index = index_cls(space=manifest["space"], ef_construction=manifest["ef_construction"], M=manifest["M"],
                      rebuild_threshold=manifest["rebuild_threshold"], device=device,
                      auto_compact=manifest["auto_compact"])
    pre = f"g{int(manifest['generation'])}." if "generation" in manifest else ""
    for i, m in enumerate(manifest["namespaces"]):
        n = int(m["rows"])
        ns = index._get_or_create(m["name"], int(m["dim"]), m["space"], capacity=n)
        rows = np.load(os.path.join(path, f"{pre}ns{i}.rows.npy"), mmap_mode="r")
        live = np.load(os.path.join(path, f"{pre}ns{i}.live.npy"))
        if rows.shape != (n, ns.dim) or live.shape[0] < (n + 31) // 32:
            raise ValueError(f"snapshot namespace {m['name']!r}: files do not match the manifest")
        for r0 in range(0, n, CHUNK_ROWS):
            nr = min(CHUNK_ROWS, n - r0)
            first = ns.shard.import_rows(np.ascontiguousarray(rows[r0:r0 + nr]), live[r0 // 32: (r0 + nr + 31) // 32])
            assert first == r0
        ns.reserve(n)
        ns.ids[:n] = np.load(os.path.join(path, f"{pre}ns{i}.ids.npy"))
        bits = np.unpackbits(live.view(np.uint8), bitorder="little")[:n].astype(bool)
        ns.gone[:n] = ~bits
        ns.n = n
        ns.total, ns.deleted = int(m["total"]), int(m["deleted"])
        ns.rebuild_required = bool(m["rebuild_required"])
        ns.uuid_to_row = None                      # rebuilt lazily from ns.ids by the first remove()
        ns.codec = ColumnCodec.from_json(m["codec"])
        for j in m["columns"]:
            ns.shard.set_column(int(j), np.load(os.path.join(path, f"{pre}ns{i}.col{j}.npy")), 0)
        ns.touch()
    return index
