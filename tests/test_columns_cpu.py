"""CPU: host logic of the columnar metadata store (``mlvectordb_b200/columns.py``).  The codec's device
predicates, evaluated by the oracle's ``where_mask`` (the statement of ``mlv_filter_create_where``), must
decide every constraint exactly as ``host_predicate`` decides it against the metadata mappings -- the
``metadata.get(key) == value`` reading of the reference's filter sketch (``README.md:477``,
``examples/api_client.py:65-74``)."""
import random

import numpy as np
import pytest

from mlvectordb_b200 import _capi
from mlvectordb_b200.columns import ColumnCodec, host_predicate
from oracle.exact import COLUMN_MISSING, where_mask


def _ingest(codec, mds, blocks):
    n = len(mds)
    cols, at = {}, 0
    for size in blocks:
        blk = mds[at:at + size]
        for c, a in codec.encode_rows(blk).items():
            assert a.dtype == np.int32 and a.shape == (len(blk),)
            cols.setdefault(c, np.full(n, COLUMN_MISSING, dtype=np.int64))[at:at + len(blk)] = a
        at += len(blk)
    assert at == n
    return cols


def _random_metadata(n, seed):
    rnd = random.Random(seed)
    mds = []
    for _ in range(n):
        md = {}
        if rnd.random() < 0.9:
            md["bucket"] = rnd.randint(-5, 20)
        if rnd.random() < 0.8:
            md["color"] = rnd.choice(["red", "green", "blue", 1, 1.0, True, None, (1, 2)])
        if rnd.random() < 0.5:
            md["w"] = rnd.choice([1.5, 2.5, "x"])
        mds.append(md if md or rnd.random() < 0.5 else None)
    return mds


CONSTRAINTS = [
    {"bucket": 3}, {"bucket": 3.0}, {"bucket": True}, {"bucket": 3.5}, {"bucket": "3"}, {"bucket": ("!=", 3)},
    {"bucket": ("<", 4)}, {"bucket": ("<", 3.5)}, {"bucket": ("<=", -5)}, {"bucket": (">", 19)}, {"bucket": (">=", 2.5)},
    {"bucket": ("between", -2, 7.5)}, {"bucket": ("between", 7, 2)}, {"bucket": (">", 2 ** 40)}, {"bucket": ("<=", -2 ** 40)},
    {"bucket": ("<", 2 ** 40)}, {"bucket": ("!=", 2 ** 40)}, {"bucket": ("!=", 0.5)},
    {"color": "red"}, {"color": 1}, {"color": ("!=", "red")}, {"color": "purple"}, {"color": ("!=", "purple")}, {"color": (1, 2)},
    {"bucket": 3, "color": "blue"}, {"bucket": (">=", 0), "color": ("!=", "green"), "w": 1.5}, {"w": "x"}, {},
]


@pytest.mark.parametrize("blocks", [(500,), (100, 250, 150), (1,) * 40 + (460,)])
def test_device_predicates_decide_like_the_host_predicate(blocks):
    mds = _random_metadata(500, seed=len(blocks))
    codec = ColumnCodec()
    cols = _ingest(codec, mds, blocks)
    assert codec.kind("bucket") == "raw" and codec.kind("color") == "dict"
    for cons in CONSTRAINTS:
        preds = codec.predicates(cons)
        assert preds is not None, cons
        assert len(preds) <= _capi.MAX_PREDICATES
        for p in preds:
            assert 0 <= p[0] < _capi.MAX_COLUMNS and p[1] in _capi.PRED_OPS
            assert -(2 ** 31) < p[2] < 2 ** 31 and -(2 ** 31) < p[3] < 2 ** 31
        want = np.array([host_predicate(cons)(md or {}) for md in mds])
        got = where_mask(cols, preds, len(mds))
        assert np.array_equal(got, want), (cons, preds, int(got.sum()), int(want.sum()))


def test_what_the_device_cannot_decide_is_reported():
    codec = ColumnCodec()
    _ingest(codec, [{"a": 1, "tags": ["x"], "s": "u"}, {"a": 2, "tags": ["y"], "s": "v"}], (2,))
    assert codec.kind("tags") == "host"                     # unhashable values
    assert codec.predicates({"tags": ["x"]}) is None
    assert codec.predicates({"a": None}) is None            # metadata.get(k) == None also matches missing keys
    assert codec.predicates({"never": 1}) is None           # rows may have been loaded without their metadata
    assert codec.predicates({"s": ("<", "v")}) is not None  # "u" < "v" arrived in order: the codes already ascend
    codec.encode_rows([{"s": "a"}])                         # ... a value that breaks the order: not decidable until re-coded
    assert codec.predicates({"s": ("<", "v")}) is None and codec.unordered_columns({"s": ("<", "v")}) == ["s"]
    mixed = ColumnCodec()
    _ingest(mixed, [{"m": "x"}, {"m": 3}, {"m": float("nan")}], (3,))
    assert mixed.unordered_columns({"m": (">", 1)}) == ["m"] and mixed.reorder("m") is None   # strings, numbers and NaN do not order
    assert mixed.predicates({"m": (">", 1)}) is None and mixed.predicates({"m": "x"}) is not None
    assert codec.predicates({"a": 1}) is not None
    # a raw column that later meets a non-integer value is given up (and says so)
    codec.encode_rows([{"a": "three"}])
    assert codec.kind("a") == "host" and codec.predicates({"a": 1}) is None
    # only MAX_COLUMNS keys live on the device
    many = ColumnCodec()
    many.encode_rows([{f"k{i}": i for i in range(_capi.MAX_COLUMNS + 3)}])
    assert len(many.names()) == _capi.MAX_COLUMNS
    assert many.predicates({f"k{_capi.MAX_COLUMNS + 1}": 1}) is None
    assert many.predicates({f"k{i}": i for i in range(_capi.MAX_PREDICATES + 1)}) is None


def test_whole_column_ingest_and_codec_round_trip():
    codec = ColumnCodec()
    idx, codes = codec.encode_column("bucket", np.arange(-3, 50, dtype=np.int64))
    assert codes.dtype == np.int32 and np.array_equal(codes, np.arange(-3, 50))
    jdx, names = codec.encode_column("name", np.array(["b", "a", "b", "c"]))
    assert jdx != idx and names[0] == names[2] and len(set(names.tolist())) == 3
    kdx, mixed = codec.encode_column("mixed", ["x", 2.5, ("t", 1), None, "x"])
    assert mixed[0] == mixed[4] and len(set(mixed.tolist())) == 4
    _, big = codec.encode_column("big", np.array([2 ** 40, 1]))      # out of int32 range: dictionary coded (codes by rank here)
    assert codec.kind("big") == "dict" and big[0] != big[1]
    lt5 = codec.predicates({"big": ("<", 5)})                        # ordered constraint = a range of codes
    assert lt5 is not None and where_mask({lt5[0][0]: big}, lt5, 2).tolist() == [False, True]
    codec.encode_column("small", np.array([1, 2]))
    assert codec.encode_column("small", np.array([2 ** 40])) is None and codec.kind("small") == "host"
    again = ColumnCodec.from_json(__import__("json").loads(__import__("json").dumps(codec.to_json())))
    for cons in ({"name": "a"}, {"name": ("!=", "zz")}, {"bucket": ("between", 0, 9)}, {"mixed": 2.5}, {"mixed": ("t", 1)},
                 {"mixed": "nope"}):
        assert again.predicates(cons) == codec.predicates(cons), cons
    assert again.predicates({"big": 2 ** 40}) == codec.predicates({"big": 2 ** 40})
    assert again.predicates({"small": 1}) is None
    # new values keep extending the restored dictionary without colliding with old codes
    _, more = again.encode_column("name", np.array(["c", "d"]))
    assert more[0] == names[3] and more[1] not in names.tolist()


def _apply(cols, column, perm):
    """what GpuIndex._order_columns does on the device column: codes -> ranks"""
    codes = cols[column]
    has = codes != COLUMN_MISSING
    codes[has] = perm[codes[has]]


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_ordered_comparisons_on_dictionary_columns_after_recoding(seed):
    """VERDICT r1 missing #5: ``{"name": ("<", "m")}`` on a string (dictionary-coded) key.  The codec re-codes the column
    by the rank of its values; afterwards every ordered constraint is a range of codes and decides exactly like the host
    predicate -- also for bounds that are not in the dictionary, bounds of another type, and values added later."""
    rnd = random.Random(seed)
    words = ["pear", "apple", "fig", "kiwi", "plum", "date", "lime", "yuzu", "nut", "apricot"]
    mds = [({"name": rnd.choice(words), "score": rnd.choice([0.5, 1.25, 2.0, 7.5, -3.0])} if rnd.random() < 0.9 else {}) for _ in range(400)]
    codec = ColumnCodec()
    cols = _ingest(codec, mds, (150, 250))
    assert codec.kind("name") == "dict" and codec.kind("score") == "dict"
    cases = [{"name": ("<", "kiwi")}, {"name": ("<=", "kiwi")}, {"name": (">", "kiwi")}, {"name": (">=", "kz")}, {"name": ("<", "a")},
             {"name": (">", "zzz")}, {"name": ("between", "b", "m")}, {"name": ("between", "m", "b")}, {"name": ("<", 3)},
             {"score": (">", 1)}, {"score": ("<=", 1.25)}, {"score": ("between", 0, 2)}, {"score": (">=", "x")},
             {"name": (">=", "fig"), "score": ("<", 2.0)}, {"name": "fig", "score": (">", 0)}]

    def check():
        for cons in cases:
            for name in codec.unordered_columns(cons):
                perm = codec.reorder(name)
                assert perm is not None
                _apply(cols, codec.column_index(name), perm)
            preds = codec.predicates(cons)
            assert preds is not None, cons
            want = np.array([host_predicate(cons)(md or {}) for md in mds])
            got = where_mask(cols, preds, len(mds))
            assert np.array_equal(got, want), (cons, preds, int(got.sum()), int(want.sum()))

    check()
    assert codec.unordered_columns({"name": ("<", "x")}) == []           # re-coded once, stays ordered
    # equality still works on the re-coded column, and a snapshot keeps the order
    back = ColumnCodec.from_json(codec.to_json())
    assert back.predicates({"name": ("<", "kiwi")}) == codec.predicates({"name": ("<", "kiwi")})
    # new values arrive out of order: the next ordered constraint re-codes again
    extra = [{"name": "banana", "score": 0.75}, {"name": "zebra"}, {"name": "cherry", "score": 100.0}]
    n0 = len(mds)
    mds.extend(extra)
    for c, a in codec.encode_rows(extra).items():
        grown = np.full(len(mds), COLUMN_MISSING, dtype=np.int64)
        grown[:n0] = cols[c]
        grown[n0:] = a
        cols[c] = grown
    assert codec.unordered_columns({"name": ("<", "x")}) == ["name"]
    check()
