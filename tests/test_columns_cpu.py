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
    assert codec.predicates({"s": ("<", "v")}) is None      # dictionary codes carry no order
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
    _, big = codec.encode_column("big", np.array([2 ** 40, 1]))      # out of int32 range: dictionary coded, equality only
    assert codec.kind("big") == "dict" and big[0] != big[1] and codec.predicates({"big": ("<", 5)}) is None
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
