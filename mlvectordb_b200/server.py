"""``python -m mlvectordb_b200.server``: the reference's server entry point (``src/mlvectordb/api/server.py:15-72``:
same ``--host --port --reload --log-level`` flags) wired to the GPU index -- the one-line change of SURVEY.md
section 2: ``QueryProcessor(StorageEngineInMemory(), Index())`` (``server.py:54``) becomes
``GpuQueryProcessor(storage, GpuIndex(space, device))``.  The storage engine stays the caller's: pass any object
implementing the reference ``StorageEngine`` protocol (``interfaces/storage_engine.py:16-53``) to ``build_app``;
run from the reference's repository root the default is its own ``StorageEngineInMemory``."""
from __future__ import annotations

import argparse


def build_app(storage_engine, space: str = "l2", device: int = 0, log_level: str = "INFO", devices=None, **index_kw):
    """``devices``: several GPUs behind the one process (``MultiGpuIndex``); otherwise one ``GpuIndex`` on ``device``."""
    from .index import GpuIndex
    from .multi import MultiGpuIndex
    from .query_processor import GpuQueryProcessor
    from .rest_api import GpuRestAPI
    index = MultiGpuIndex(space=space, devices=list(devices), **index_kw) if devices else GpuIndex(space=space, device=device, **index_kw)
    processor = GpuQueryProcessor(storage_engine, index)
    return GpuRestAPI(processor, title="Vector DB API (B200 exact index)", log_level=log_level).get_app()


def main(argv=None) -> None:
    ap = argparse.ArgumentParser(description="MLVectorDB REST server over the B200 exact-search index")
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--port", type=int, default=8000)
    ap.add_argument("--reload", action="store_true")
    ap.add_argument("--log-level", default="info", choices=["debug", "info", "warning", "error"])
    ap.add_argument("--space", default="l2", help="hnswlib space of the index: l2 | ip | cosine (reference default: l2)")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--devices", default="", help="comma-separated GPUs to spread every namespace over (MultiGpuIndex), e.g. 0,1,2,3")
    a = ap.parse_args(argv)
    import uvicorn
    try:
        from src.mlvectordb.implementations.storage_engine_in_memory import StorageEngineInMemory  # reference checkout
    except Exception as e:  # noqa: BLE001
        raise SystemExit("no storage engine: run from the reference's repository root (its StorageEngineInMemory is "
                         f"reused unchanged) or call build_app(storage_engine) yourself ({e})")
    devices = [int(d) for d in a.devices.split(",") if d.strip() != ""]
    uvicorn.run(build_app(StorageEngineInMemory(), a.space, a.device, a.log_level.upper(), devices=devices), host=a.host, port=a.port,
                reload=a.reload, log_config=None)


if __name__ == "__main__":
    main()
