"""OPQ pre-transform: "OPQ16,IVF...,PQ16" keys of the reference's recall studies
(bench_cpu_recall.py:54).  The rotation is build-time library math; what is checked here is that it is a rotation,
that it does what OPQ is for (lower PQ error on correlated data), and that the search behind it is still the oracle's
search on the rotated inputs, bit for bit."""
import numpy as np
import pytest

import _util

pytestmark = pytest.mark.gpu


def _correlated(n, d, seed):
    rng = np.random.default_rng(seed)
    mix = rng.standard_normal((d, d)).astype(np.float32) / np.sqrt(d)
    scales = np.exp(np.linspace(1.5, -1.5, d)).astype(np.float32)           # strongly anisotropic
    centers = rng.standard_normal((64, d)).astype(np.float32) * 2
    x = centers[rng.integers(0, 64, n)] + (rng.standard_normal((n, d)).astype(np.float32) * scales) @ mix
    return np.ascontiguousarray(x, np.float32)


def test_opq_rotation_and_search_parity(oracle):
    import torch
    import b200ivfpq as faiss
    d, nlist, M, nb, nq = 64, 32, 16, 20000, 64
    xb, xq = _correlated(nb, d, 0), _correlated(nq, d, 1)
    index = faiss.index_factory(d, f"OPQ{M},IVF{nlist},PQ{M}")
    index.chain.at(0).niter = 12
    index.train(xb)
    assert index.is_trained
    vt = faiss.downcast_VectorTransform(index.chain.at(0))
    A = faiss.vector_to_array(vt.A).reshape(d, d)                           # extract_FPGA_required_data.py:165-167
    np.testing.assert_allclose(A @ A.T, np.eye(d), atol=2e-4)
    index.add(xb)
    index.nprobe = 8
    D, I = index.search(xq, 10)
    assert D.shape == (nq, 10) and I.dtype == np.int64 and index.ntotal == nb
    # parity: the oracle on the sub-index's arrays and the rotated queries
    sub = faiss.downcast_index(index.index)
    a = sub.to_arrays()
    xr = vt.apply(xq)
    Dr, Ir = oracle.C.search(np.ascontiguousarray(xr), a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], 8, 10)
    _util.assert_bit_equal(D, Dr, "D behind OPQ")
    _util.assert_bit_equal(I, Ir, "I behind OPQ")
    # OPQ's purpose: lower reconstruction error than the same PQ without the rotation
    plain = faiss.index_factory(d, f"IVF{nlist},PQ{M}")
    plain.train(xb)

    def pq_error(ix, x):
        xt = torch.from_numpy(x).cuda()
        cent, pq = ix.quantizer.xb_tensor(), ix.pq.centroids_tensor()
        lab = torch.cdist(xt, cent).argmin(1)
        res = (xt - cent[lab]).reshape(-1, M, d // M)
        dist = torch.cdist(res.permute(1, 0, 2), pq)                        # (M, n, 256)
        return float((dist.min(2).values ** 2).sum() / x.shape[0])

    e_opq, e_plain = pq_error(sub, vt.apply(xb[:5000])), pq_error(plain, xb[:5000])
    assert e_opq < 0.98 * e_plain, (e_opq, e_plain)


def test_opq_dimension_reduction():
    import b200ivfpq as faiss
    xb = _correlated(6000, 64, 2)
    index = faiss.index_factory(64, "OPQ8_32,IVF8,PQ8")
    index.chain.at(0).niter = 4
    index.train(xb)
    index.add(xb)
    index.nprobe = 8
    A = index.chain.at(0).A.reshape(32, 64)
    np.testing.assert_allclose(A @ A.T, np.eye(32), atol=2e-4)
    D, I = index.search(xb[:20], 5)
    assert (I[:, 0] >= 0).all() and (np.diff(D, axis=1) >= 0).all()
    hit = (I[:, :5] == np.arange(20)[:, None]).any(1).mean()               # a vector should find itself
    assert hit >= 0.7
