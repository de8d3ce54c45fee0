"""Host side of the columnar metadata store.

The reference keeps metadata as one arbitrary mapping per vector (``implementations/vector.py:15``,
``interfaces/vector.py:8-22``) and only sketches filters: a dict of equality constraints
(``README.md:123,477``; request shape ``examples/api_client.py:65-74``).  The device can only compare
int32 codes (``include/mlv_index.h``: ``mlv_index_set_column`` / ``mlv_filter_create_where``), so this
module turns metadata keys into columns and values into codes:

* a key whose values are Python ``int`` (not ``bool``) inside the int32 range is stored **raw**, so
  ordered comparisons (``< <= > >= between``) work on it;
* any other hashable value is **dictionary coded** (value -> 0, 1, 2 ... in order of first appearance;
  Python equality decides: ``1 == 1.0 == True`` share a code).  ``==`` / ``!=`` compare codes; an ORDERED comparison
  (``< <= > >= between``) needs order-preserving codes: when the dictionary's values are mutually comparable the index
  re-codes the column by rank once (``reorder`` -> one column rewrite on the device) and the comparison becomes a
  range of codes; values that cannot be ordered (mixed strings and numbers, NaN) keep the host fallback;
* a key that cannot be represented (unhashable values, a raw column that meets a non-int value, more than
  ``MAX_COLUMNS`` keys) is marked *host only* and constraints on it are reported as not device-evaluable,
  so the caller falls back to evaluating the predicate against the stored metadata.

A row without the key holds ``COLUMN_MISSING`` and fails every constraint on that key -- the same answer
``metadata.get(key) == value`` gives for every value except ``None`` (constraints with ``None`` are
reported as not device-evaluable for that reason).
"""
from __future__ import annotations

from typing import Any, Dict, List, Mapping, Optional, Sequence, Tuple

import numpy as np

from ._capi import COLUMN_MISSING, MAX_COLUMNS, MAX_PREDICATES, PRED_OPS

_I32_MIN, _I32_MAX = -(2 ** 31) + 1, 2 ** 31 - 1   # COLUMN_MISSING itself is not a storable value
_IMPOSSIBLE = ("between", 1, 0)                     # a predicate no row satisfies

Predicate = Tuple[int, str, int, int]


def _is_raw_int(v: Any) -> bool:
    return isinstance(v, (int, np.integer)) and not isinstance(v, (bool, np.bool_)) and _I32_MIN <= int(v) <= _I32_MAX


def _tag(v: Any):
    if v is None or isinstance(v, (bool, str)):
        return v
    if isinstance(v, (int, np.integer)):
        return int(v)
    if isinstance(v, (float, np.floating)):
        return {"f": float(v).hex()}
    if isinstance(v, tuple):
        return {"t": [_tag(x) for x in v]}
    raise TypeError(type(v))


def _untag(t: Any):
    if isinstance(t, dict):
        return float.fromhex(t["f"]) if "f" in t else tuple(_untag(x) for x in t["t"])
    return t


class _Column:
    __slots__ = ("index", "kind", "codes", "ordered", "top")

    def __init__(self, index: int, kind: str):
        self.index = index
        self.kind = kind                  # "raw" | "dict" | "host"
        self.codes: Dict[Any, int] = {}   # dict kind: value -> code
        self.ordered = True               # dict kind: codes ascend with the values (Python <), so code ranges are value ranges
        self.top: Any = None              # dict kind, ordered: the value holding the largest code


def _orderable(v: Any) -> bool:
    return not (isinstance(v, (float, np.floating)) and v != v)   # NaN orders with nothing


class ColumnCodec:
    """Per-namespace mapping metadata key -> device column, value -> int32 code."""

    def __init__(self):
        self._cols: Dict[str, _Column] = {}
        self._n_device = 0

    # ------------------------------------------------------------------ introspection
    def column_index(self, name: str) -> Optional[int]:
        c = self._cols.get(name)
        return c.index if c is not None and c.kind != "host" else None

    def kind(self, name: str) -> Optional[str]:
        c = self._cols.get(name)
        return c.kind if c is not None else None

    def names(self) -> List[str]:
        return [n for n, c in self._cols.items() if c.kind != "host"]

    # ------------------------------------------------------------------ ingest
    def _column_for(self, name: str, first_value: Any) -> _Column:
        c = self._cols.get(name)
        if c is None:
            if self._n_device >= MAX_COLUMNS:
                c = _Column(-1, "host")
            else:
                c = _Column(self._n_device, "raw" if _is_raw_int(first_value) else "dict")
                self._n_device += 1
            self._cols[name] = c
        return c

    def _code(self, c: _Column, v: Any) -> Optional[int]:
        """int32 code of ``v`` in column ``c`` or None when ``c`` cannot hold it."""
        if c.kind == "raw":
            if _is_raw_int(v):
                return int(v)
            if isinstance(v, (bool, float, np.floating, np.bool_)) and v == int(v) and _I32_MIN <= int(v) <= _I32_MAX:
                return int(v)            # 1.0 / True == 1, exactly as Python compares them
            return None
        try:
            code = c.codes.get(v)
            if code is None:
                code = len(c.codes)
                if code > _I32_MAX:
                    return None
                if c.ordered and code:      # does the new value extend the order?  (mixed types / NaN: it does not)
                    try:
                        c.ordered = bool(_orderable(v) and c.top < v)
                    except TypeError:
                        c.ordered = False
                elif not code:
                    c.ordered = _orderable(v)
                c.codes[v] = code
                if c.ordered:
                    c.top = v
            return code
        except TypeError:                # unhashable
            return None

    def encode_rows(self, metadatas: Sequence[Optional[Mapping[str, Any]]]) -> Dict[int, np.ndarray]:
        """Per-row metadata mappings -> {device column: int32[len(metadatas)]} for every column that holds a value
        in this block (``COLUMN_MISSING`` where a row lacks the key)."""
        n = len(metadatas)
        out: Dict[int, np.ndarray] = {}
        for i, md in enumerate(metadatas):
            if not md:
                continue
            for name, v in md.items():
                c = self._column_for(name, v)
                if c.kind == "host":
                    continue
                code = self._code(c, v)
                if code is None:         # the column cannot represent this value: give the key up
                    c.kind = "host"
                    out.pop(c.index, None)
                    continue
                arr = out.get(c.index)
                if arr is None:
                    arr = out[c.index] = np.full(n, COLUMN_MISSING, dtype=np.int32)
                arr[i] = code
        return {idx: a for idx, a in out.items() if self._kind_of_index(idx) != "host"}

    def _kind_of_index(self, idx: int) -> str:
        for c in self._cols.values():
            if c.index == idx:
                return c.kind
        return "host"

    def encode_column(self, name: str, values) -> Optional[Tuple[int, np.ndarray]]:
        """A whole column at once (bulk ingest): integer arrays are stored raw, anything else dictionary coded.
        -> (device column, int32 codes) or None when the key is host only."""
        if isinstance(values, np.ndarray):
            arr = values.reshape(-1)
        else:
            values = list(values)
            arr = np.empty(len(values), dtype=object)
            arr[:] = values
            if values and all(_is_raw_int(v) for v in values):
                arr = arr.astype(np.int64)
        first = arr[0] if arr.size else 0
        first = first.item() if isinstance(first, np.generic) else first
        c = self._column_for(name, first)
        if c.kind == "host":
            return None
        if c.kind == "raw":
            if arr.dtype.kind in "iu" and (arr.size == 0 or (arr.min() >= _I32_MIN and arr.max() <= _I32_MAX)):
                return c.index, arr.astype(np.int32)
            codes = [self._code(c, v) for v in arr.tolist()]
        else:
            uniq, inverse = np.unique(arr, return_inverse=True) if arr.dtype.kind in "iuUSfb" else (None, None)
            if uniq is not None:
                lut = np.empty(len(uniq), dtype=np.int64)
                for j, u in enumerate(uniq.tolist()):
                    code = self._code(c, u)
                    if code is None:
                        c.kind = "host"
                        return None
                    lut[j] = code
                return c.index, lut[inverse.reshape(-1)].astype(np.int32)
            codes = [self._code(c, v) for v in arr.tolist()]
        if any(code is None for code in codes):
            c.kind = "host"
            return None
        return c.index, np.asarray(codes, dtype=np.int32)

    # ------------------------------------------------------------------ ordered comparisons on dictionary columns
    def unordered_columns(self, constraints: Mapping[str, Any]) -> List[str]:
        """Dictionary columns that an ordered constraint of ``constraints`` touches while their codes are not in value
        order: ``reorder`` them (and rewrite the device column) before asking for ``predicates``."""
        out = []
        for name, want in constraints.items():
            if isinstance(want, tuple) and len(want) in (2, 3) and want[0] in PRED_OPS and want[0] not in ("==", "!="):
                c = self._cols.get(name)
                if c is not None and c.kind == "dict" and not c.ordered and len(c.codes) > 1:
                    out.append(name)
        return out

    def reorder(self, name: str) -> Optional[np.ndarray]:
        """Re-code dictionary column ``name`` by the rank of its values.  -> int32 array old code -> new code (the caller
        rewrites the stored column through it), or None when the values cannot be ordered (they stay as they are and
        ordered constraints on the column remain host-evaluated)."""
        c = self._cols.get(name)
        if c is None or c.kind != "dict":
            return None
        values = list(c.codes)
        try:
            if not all(_orderable(v) for v in values):
                return None
            ranked = sorted(values)
            if any(not (a < b) for a, b in zip(ranked, ranked[1:])):     # a total, strict order or nothing
                return None
        except TypeError:
            return None
        perm = np.empty(len(values), dtype=np.int32)
        for new, v in enumerate(ranked):
            perm[c.codes[v]] = new
        c.codes = {v: i for i, v in enumerate(ranked)}
        c.ordered, c.top = True, (ranked[-1] if ranked else None)
        return perm

    # ------------------------------------------------------------------ snapshot (snapshot.py)
    def to_json(self) -> dict:
        """JSON-able state.  Dictionary codes survive for str / int / float / bool / None values and tuples of
        them; a column holding anything else is written as host only (its device column is not saved)."""
        cols = {}
        for name, c in self._cols.items():
            kind, codes = c.kind, []
            if kind == "dict":
                try:
                    codes = [[_tag(v), code] for v, code in c.codes.items()]
                except TypeError:
                    kind, codes = "host", []
            cols[name] = {"index": c.index if kind != "host" else -1, "kind": kind, "codes": codes}
        return {"columns": cols, "n_device": self._n_device}

    @classmethod
    def from_json(cls, state: dict) -> "ColumnCodec":
        self = cls()
        self._n_device = int(state["n_device"])
        for name, st in state["columns"].items():
            c = _Column(int(st["index"]), st["kind"])
            c.codes = {_untag(t): int(code) for t, code in st["codes"]}
            by_code = sorted(c.codes, key=c.codes.get)
            try:
                c.ordered = all(_orderable(v) for v in by_code) and all(a < b for a, b in zip(by_code, by_code[1:]))
            except TypeError:
                c.ordered = False
            c.top = by_code[-1] if (c.ordered and by_code) else None
            self._cols[name] = c
        return self

    # ------------------------------------------------------------------ constraints -> device predicates
    def predicates(self, constraints: Mapping[str, Any]) -> Optional[List[Predicate]]:
        """``{key: value}`` (equality) or ``{key: (op, a)}`` / ``{key: ("between", a, b)}`` ->
        ``[(column, op, a, b)]`` for ``DeviceShard.where``; None when the device cannot evaluate it."""
        preds: List[Predicate] = []
        for name, want in constraints.items():
            if isinstance(want, tuple) and len(want) in (2, 3) and want[0] in PRED_OPS:
                op, args = want[0], want[1:]
            else:
                op, args = "==", (want,)
            if any(a is None for a in args):
                return None              # metadata.get(key) == None also matches rows WITHOUT the key
            c = self._cols.get(name)
            if c is None or c.kind == "host":
                return None              # never seen here (rows may have been loaded without their metadata) / host only
            if c.kind == "dict" and op not in ("==", "!="):
                if not c.ordered:
                    return None          # codes carry no order (yet: see unordered_columns / reorder)
                pred = _dict_range_predicate(c, op, args)
                if pred is None:
                    return None
                preds.append(pred)
                continue
            if c.kind == "dict":
                try:
                    code = c.codes.get(args[0])
                except TypeError:
                    return None
                if code is None:         # value never seen: nothing equals it, everything present differs
                    preds.append((c.index, ">=", 0, 0) if op == "!=" else (c.index,) + _IMPOSSIBLE)
                else:
                    preds.append((c.index, op, code, 0))
                continue
            vals = []
            for a in args:
                if isinstance(a, (int, float, bool, np.integer, np.floating, np.bool_)):
                    vals.append(a)
                else:
                    vals.append(None)
            if any(v is None for v in vals):
                if op == "==":
                    preds.append((c.index,) + _IMPOSSIBLE)   # an int never equals a string / tuple ...
                    continue
                return None
            pred = _raw_predicate(c.index, op, vals)
            if pred is None:
                return None
            preds.append(pred)
        if len(preds) > MAX_PREDICATES:
            return None
        return preds


def _dict_range_predicate(c: _Column, op: str, args) -> Optional[Predicate]:
    """Ordered comparison on a dictionary column whose codes ascend with its values: a range of codes.  A bound that
    does not compare with the values (``"a" < 3``) satisfies nothing, exactly as ``host_predicate`` decides it."""
    import bisect
    vals = list(c.codes)      # insertion order == code order == value order
    try:
        if not all(_orderable(a) for a in args):
            return (c.index,) + _IMPOSSIBLE
        if op == "between":
            lo, hi = bisect.bisect_left(vals, args[0]), bisect.bisect_right(vals, args[1]) - 1
            return (c.index, "between", lo, hi) if lo <= hi else (c.index,) + _IMPOSSIBLE
        if op == "<":
            return (c.index, "<", bisect.bisect_left(vals, args[0]), 0)
        if op == "<=":
            return (c.index, "<", bisect.bisect_right(vals, args[0]), 0)
        if op == ">":
            return (c.index, ">=", bisect.bisect_right(vals, args[0]), 0)
        return (c.index, ">=", bisect.bisect_left(vals, args[0]), 0)
    except TypeError:
        return (c.index,) + _IMPOSSIBLE


def _raw_predicate(column: int, op: str, vals) -> Optional[Predicate]:
    """Comparison of a raw int column with possibly fractional / out-of-range numbers, as Python would decide it."""
    import math
    a = vals[0]
    if isinstance(a, float) and (math.isnan(a)):
        return (column, ">=", _I32_MIN, 0) if op == "!=" else (column,) + _IMPOSSIBLE
    if op in ("==", "!="):
        exact = (not isinstance(a, float) or a == math.floor(a)) and _I32_MIN <= a <= _I32_MAX
        if not exact:
            return (column, ">=", _I32_MIN, 0) if op == "!=" else (column,) + _IMPOSSIBLE
        return (column, op, int(a), 0)
    if op == "between":
        b = vals[1]
        if isinstance(b, float) and math.isnan(b):
            return (column,) + _IMPOSSIBLE
        lo, hi = math.ceil(a), math.floor(b)
        lo, hi = max(lo, _I32_MIN), min(hi, _I32_MAX)
        return (column, "between", int(lo), int(hi)) if lo <= hi else (column,) + _IMPOSSIBLE
    # ordered comparison with a bound that may be fractional or outside int32
    if op in ("<", "<="):
        bound = math.ceil(a) - 1 if op == "<" else math.floor(a)       # value <= bound
        if bound < _I32_MIN:
            return (column,) + _IMPOSSIBLE
        return (column, "<=", int(min(bound, _I32_MAX)), 0)
    bound = math.floor(a) + 1 if op == ">" else math.ceil(a)           # value >= bound
    if bound > _I32_MAX:
        return (column,) + _IMPOSSIBLE
    return (column, ">=", int(max(bound, _I32_MIN)), 0)


_HOST_OPS = {
    "==": lambda v, a: v == a[0],
    "!=": lambda v, a: v != a[0],
    "<": lambda v, a: v < a[0],
    "<=": lambda v, a: v <= a[0],
    ">": lambda v, a: v > a[0],
    ">=": lambda v, a: v >= a[0],
    "between": lambda v, a: a[0] <= v <= a[1],
}


def host_predicate(constraints: Mapping[str, Any]):
    """The same constraints decided against one metadata mapping on the host (the fallback when the device
    columns cannot decide them, and the statement of what the device path must return): equality is
    ``metadata.get(key) == value``; an operator constraint needs the key to be present and comparable."""
    items = []
    for name, want in constraints.items():
        if isinstance(want, tuple) and len(want) in (2, 3) and want[0] in PRED_OPS:
            items.append((name, want[0], want[1:]))
        else:
            items.append((name, None, want))

    def pred(md: Mapping[str, Any]) -> bool:
        for name, op, arg in items:
            if op is None:
                if md.get(name) != arg:
                    return False
                continue
            if name not in md:
                return False
            try:
                if not _HOST_OPS[op](md[name], arg):
                    return False
            except TypeError:
                return False
        return True

    return pred
