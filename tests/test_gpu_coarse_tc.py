"""K1 on tensor cores: the tcgen05 split-bf16 GEMM scores, the error bound used by the sufficiency proof, and the
end result (probed lists + distances identical to the oracle, with and without the tensor-core path)."""
import ctypes
import os

import numpy as np
import pytest

import _util

pytestmark = pytest.mark.gpu


def _flat_handle(cent):
    import torch
    import b200ivfpq as faiss
    idx = faiss.IndexFlatL2(cent.shape[1])
    idx.add(cent)
    return idx, idx._ensure_handle()


@pytest.mark.parametrize("nq,nlist,d", [(300, 1000, 128), (1, 513, 96), (257, 8192, 128), (64, 2048, 768), (130, 300, 20)])
def test_tc_scores_match_fp64(nq, nlist, d):
    """s(q, c) = max(0, ||q||^2 + ||c||^2 - 2 q.c) from the tcgen05 GEMM vs float64; error must sit well inside the bound the
    sufficiency proof uses: E = 6e-5 ||q|| ||c||max + (d+2) 2^-23 (||q|| + ||c||max)^2."""
    import torch
    import b200ivfpq as faiss
    from b200ivfpq import _lib
    rng = np.random.default_rng(nq + nlist)
    cent = (rng.random((nlist, d), dtype=np.float32) * 2 - 0.5).astype(np.float32)
    xq = (rng.random((nq, d), dtype=np.float32) * 2 - 0.5).astype(np.float32)
    idx, h = _flat_handle(cent)
    xq_t = torch.from_numpy(xq).cuda()
    out = torch.empty((nq, nlist), dtype=torch.float32, device="cuda")
    _lib.check(h.lib.b200_ivfpq_coarse_scores(h.h, nq, xq_t.data_ptr(), out.data_ptr(), None))
    torch.cuda.synchronize()
    got = out.cpu().numpy().astype(np.float64)
    c64, q64 = cent.astype(np.float64), xq.astype(np.float64)
    ref = np.maximum(0.0, (q64 * q64).sum(1)[:, None] + (c64 * c64).sum(1)[None, :] - 2.0 * q64 @ c64.T)
    err = np.abs(got - ref)
    qn = np.sqrt((q64 * q64).sum(1))[:, None]
    cmax = np.sqrt((c64 * c64).sum(1)).max()
    bound = 6e-5 * qn * cmax + (d + 2) * 2.0 ** -23 * (qn + cmax) ** 2
    assert (err <= bound / 4).all(), f"max err/bound = {(err / bound).max():.3f}"


@pytest.mark.parametrize("d,nlist,nq,nprobe", [(128, 1024, 500, 16), (96, 700, 129, 64), (128, 8192, 300, 32),
                                              (768, 600, 40, 32), (64, 300, 50, 1),
                                              (128, 3001, 100, 8),       # two-pass filter, ragged last chunk
                                              (96, 65536, 150, 64),      # C3's coarse quantizer: two-pass, L = 128
                                              (32, 4096, 700, 5)])
def test_tc_coarse_identical_to_oracle(oracle, d, nlist, nq, nprobe):
    import b200ivfpq as faiss
    rng = np.random.default_rng(d + nlist)
    cent = rng.random((nlist, d), dtype=np.float32)
    xq = (cent[rng.integers(0, nlist, nq)] + rng.standard_normal((nq, d)).astype(np.float32) * 0.2).astype(np.float32)
    dr, ir = oracle.C.coarse(xq, cent, nprobe)
    for variant in ("auto", "matrix", "exact"):   # two-pass filter when nlist / 32 >= L, score matrix + select, fp32
        os.environ["B200_IVFPQ_COARSE"] = variant
        try:
            idx = faiss.IndexFlatL2(d)
            idx.add(cent)
            D, I = idx.search(xq, nprobe)
        finally:
            del os.environ["B200_IVFPQ_COARSE"]
        _util.assert_bit_equal(I, ir, f"probed ids ({variant})")
        _util.assert_bit_equal(D, dr, f"coarse distances ({variant})")
        if variant != "exact":
            from b200ivfpq import _lib
            h = idx._ensure_handle()
            n = ctypes.c_int64()
            _lib.check(h.lib.b200_ivfpq_coarse_fallbacks(h.h, ctypes.byref(n)))
            assert n.value <= max(1, nq // 50), f"{n.value} of {nq} queries fell back to the exact kernel"


@pytest.mark.parametrize("nlist", [600, 2400])   # 2400: two-pass filter, the duplicates overflow its candidate cap
def test_tc_coarse_adversarial_near_ties(oracle, nlist):
    """Centroids that differ by less than the GEMM's error: the proof must fail and the exact fallback must
    restore the oracle's answer (ties -> lower id)."""
    import b200ivfpq as faiss
    from b200ivfpq import _lib
    rng = np.random.default_rng(0)
    d, nq, nprobe = 128, 64, 8
    base = rng.random((6, d), dtype=np.float32)
    cent = base[rng.integers(0, 6, nlist)].copy()
    cent += (rng.standard_normal((nlist, d)) * 1e-6).astype(np.float32)      # 100 near-duplicates per base
    cent[10] = cent[3]                                                       # exact duplicates too
    xq = (base[rng.integers(0, 6, nq)] + rng.standard_normal((nq, d)).astype(np.float32) * 0.01).astype(np.float32)
    dr, ir = oracle.C.coarse(xq, cent, nprobe)
    idx = faiss.IndexFlatL2(d)
    idx.add(cent)
    D, I = idx.search(xq, nprobe)
    _util.assert_bit_equal(I, ir, "probed ids")
    _util.assert_bit_equal(D, dr, "coarse distances")
    h = idx._ensure_handle()
    n = ctypes.c_int64()
    _lib.check(h.lib.b200_ivfpq_coarse_fallbacks(h.h, ctypes.byref(n)))
    assert n.value > 0, "near-duplicate centroids must trigger the exact fallback"


@pytest.mark.parametrize("d,nlist,nq,nprobe", [(128, 8192, 1, 32), (768, 16384, 1, 32), (96, 1000, 5, 7), (20, 300, 16, 32),
                                              (130, 4097, 9, 1), (128, 128, 3, 32)])
def test_small_batch_coarse_identical_to_oracle(oracle, d, nlist, nq, nprobe):
    """Latency path (coarse_small.cuh): exact distances + register top-32 per 128 centroids; must equal the oracle bit
    for bit, duplicates (ties -> lower id) included."""
    import b200ivfpq as faiss
    rng = np.random.default_rng(d + nlist + nq)
    cent = rng.random((nlist, d), dtype=np.float32)
    cent[nlist // 2] = cent[3]                                   # an exact duplicate
    xq = (cent[rng.integers(0, nlist, nq)] + rng.standard_normal((nq, d)).astype(np.float32) * 0.1).astype(np.float32)
    xq[0] = cent[3]                                              # distance 0 to both duplicates
    dr, ir = oracle.C.coarse(xq, cent, nprobe)
    idx = faiss.IndexFlatL2(d)
    idx.add(cent)
    D, I = idx.search(xq, nprobe)
    _util.assert_bit_equal(I, ir, "probed ids (small batch)")
    _util.assert_bit_equal(D, dr, "coarse distances (small batch)")
