"""CPU oracle for the MLVectorDB exact-search hot path.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference's arithmetic lives in the third-party wheel
``hnswlib == 0.8.0`` (reference ``pyproject.toml:12``, ``poetry.lock:144-152``), which is
not vendored under ``/root/reference`` and is not installable here (no network, no wheel).
The reference's own tests hold no golden vectors or known-answer values for this path
(reference ``tests/test_index.py``, ``tests/test_query_processor.py`` assert only
behaviour: ordering, counts, id membership).  This package therefore *restates* hnswlib
0.8.0's published distance definitions and drives them through the reference's own
``Index`` / ``QueryProcessor`` wrappers (see ``oracle/refload.py``); the behavioural pins
of the reference tests are what it is checked against.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference
legs may import anything from here.  The product (``mlvectordb_b200``) never does.
"""
