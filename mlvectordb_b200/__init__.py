"""B200-native exact-search index for MLVectorDB's ``Index.search`` hot path.

Public surface: :class:`GpuIndex` (drop-in for the reference ``Index``), :class:`GpuQueryProcessor`
(the reference ``QueryProcessor`` plus batch / filter / range / bulk-ingest entry points), :class:`DeviceShard`
(one device-resident row matrix over the C ABI in ``include/mlv_index.h``), ``SearchResult`` /
``VectorDTO`` (value types).  There is no CPU fallback: the CUDA library must be built
(``make -C mlvectordb_b200/csrc``) and a B200 must be present to create an index.
"""
from .interfaces import IndexProtocol, SearchResult, VectorDTO, VectorProtocol
from .shard import DeviceShard, PreparedFilter, canonical_space, pack_bitmap
from .index import GpuIndex
from .multi import MultiGpuIndex
from .query_processor import GpuQueryProcessor, StoredVector

__all__ = ["GpuIndex", "MultiGpuIndex", "GpuQueryProcessor", "StoredVector", "DeviceShard", "PreparedFilter", "SearchResult", "VectorDTO",
           "VectorProtocol", "IndexProtocol", "canonical_space", "pack_bitmap"]
