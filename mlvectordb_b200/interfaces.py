"""Host-side mirror of the reference's value types for the hot path.

Same names and field meaning as the reference so callers and tests read the same:
``VectorDTO`` (reference ``src/mlvectordb/interfaces/vector.py:19-22``), ``SearchResult``
(``src/mlvectordb/implementations/index.py:11-14``), ``IndexProtocol``
(``src/mlvectordb/interfaces/index.py:9-13``).  Everything is duck-typed exactly like the
reference: the index only touches ``v.id``, ``v.values`` and ``query.values``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Iterable, List, Mapping, Protocol, Sequence, runtime_checkable
from uuid import UUID

import numpy as np


@runtime_checkable
class VectorProtocol(Protocol):
    id: UUID
    values: np.ndarray
    metadata: Mapping[str, Any]


@dataclass
class VectorDTO:
    values: Sequence[float]
    metadata: Mapping[str, Any] = field(default_factory=dict)


@dataclass
class SearchResult:
    vector_id: UUID
    score: float


class IndexProtocol(Protocol):
    def add(self, vectors: Iterable[VectorProtocol], namespace: str) -> None: ...
    def remove(self, ids: Sequence[UUID], namespace: str) -> None: ...
    def search(self, query: VectorDTO, top_k: int, namespace: str, metric: str) -> List[SearchResult]: ...
    def rebuild(self, source: Mapping[str, Iterable[VectorProtocol]], metric: str) -> None: ...
