"""GPU (>= 2 devices): ``MultiGpuIndex`` -- several GPUs behind the reference's one-process server -- returns what a
single ``GpuIndex`` returns.  Every part scans on its own device, candidates are peer-copied to the first device and
ordered by ``mlv_merge_topk``; scores are bit-identical (same scan arithmetic), ids identical."""
import numpy as np
import pytest

from oracle import synthetic

pytestmark = pytest.mark.gpu


def _n_gpus():
    from mlvectordb_b200 import _capi
    return _capi.lib().mlv_device_count()


class _V:
    def __init__(self, values, metadata):
        import uuid
        self.id, self.values, self.metadata = uuid.uuid4(), np.asarray(values, np.float32), metadata


@pytest.mark.parametrize("space", ["cosine", "l2"])
def test_multi_gpu_index_equals_single_gpu_index(space):
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    from mlvectordb_b200 import GpuIndex, MultiGpuIndex, VectorDTO
    n, dim = 30_000, 64
    X = synthetic.rows(19, 0, n, dim, scaled=True)
    vecs = [_V(X[i], {"bucket": i % 10}) for i in range(n)]
    one, many = GpuIndex(space=space, device=0), MultiGpuIndex(space=space, devices=[0, 1])
    for lo, hi in ((0, 3), (3, 10_000), (10_000, n)):
        one.add(vecs[lo:hi], "ns")
        many.add(vecs[lo:hi], "ns")
    assert many.info("ns")["rows_per_device"] == [n // 2, n // 2] and many.info("ns")["devices"] == [0, 1]
    Q = synthetic.queries(19, 48, dim)
    Q[0] = X[n - 1]

    def same(k, flt=None):
        for q in Q[:4]:
            a = one.search(VectorDTO(values=q), k, "ns", space, filter=flt)
            b = many.search(VectorDTO(values=q), k, "ns", space, filter=flt)
            assert [r.vector_id for r in a] == [r.vector_id for r in b]
            assert [r.score for r in a] == [r.score for r in b]
        return b

    top = same(10)
    assert top and same(1)[0].vector_id is not None
    assert many.search(VectorDTO(values=Q[0]), 1, "ns", space)[0].vector_id == vecs[n - 1].id
    same(200)                                              # the NVLink-merged list is longer than one warp's list
    same(5, {"bucket": 3})                                 # metadata constraint decided per part on its device columns
    allowed = {v.id for v in vecs[::13]}
    same(5, lambda u: u in allowed)                        # host predicate -> per-part bitmaps
    # batches: a scan-sized one and one that takes the tensor-core path on every part
    for nq in (3, 48):
        r1, s1, c1 = one.search_batch(Q[:nq], 10, "ns")
        r2, s2, c2 = many.search_batch(Q[:nq], 10, "ns")
        assert np.array_equal(s1, s2) and np.array_equal(c1, c2)
        assert one.uuids_of("ns", r1[nq - 1]) == many.uuids_of("ns", r2[nq - 1])
    gone = [vecs[i].id for i in range(0, n, 9)]
    one.remove(gone, "ns")
    many.remove(gone, "ns")
    same(10)
    tenth = one.search(VectorDTO(values=Q[1]), 10, "ns", space)[-1].score
    radius = float(np.float32(1 - tenth if space == "cosine" else tenth)) * (1 + 1e-6) + 1e-7
    h1 = one.range_search(VectorDTO(values=Q[1]), radius, "ns", space)
    h2 = many.range_search(VectorDTO(values=Q[1]), radius, "ns", space)
    assert [h.vector_id for h in h1] == [h.vector_id for h in h2] and [h.score for h in h1] == [h.score for h in h2] and len(h1) >= 10
    one.close()
    many.close()
