"""Row a1 against CODE OF THE REFERENCE run here: the FPGA deployment's host selects the nprobe cells with its vendored
hnswlib (BruteforceSearch + searchKnn, retrieval_accelerator/entire_accelerator_final_SIFT_M32/src/host.cpp:516-581).
oracle/ref_coarse_shim.cpp calls exactly that from the headers under /root/reference (oracle/_ref/libref_coarse.so;
built here, travels prebuilt to the GPU box).  hnswlib sums with SIMD lanes, the contract sequentially: ids must agree
except for near ties, distances to rounding."""
import numpy as np
import pytest

import _util

RTOL = 1e-5       # north_star's distance tolerance

CASES = [  # d, nlist, nq, nprobe
    (128, 1024, 200, 16),    # C1's coarse quantizer
    (96, 4096, 64, 64),      # C3's d and nprobe
    (768, 512, 32, 32),      # C4's d
    (20, 64, 100, 8),
    (128, 300, 50, 300),     # nprobe = nlist: a full ranking
]


def _data(seed, d, nlist, nq, clustered):
    rng = np.random.default_rng(seed)
    cent = rng.random((nlist, d), dtype=np.float32)
    if clustered:   # queries next to centroids: small distances, large cancellation in the GEMM forms
        xq = cent[rng.integers(0, nlist, nq)] + rng.standard_normal((nq, d)).astype(np.float32) * np.float32(0.01)
    else:
        xq = rng.random((nq, d), dtype=np.float32)
    return cent, np.ascontiguousarray(xq, np.float32)


@pytest.fixture(scope="module")
def ref(oracle):
    if oracle.build_ref() is None:
        pytest.skip("oracle/_ref/libref_coarse.so not available (reference not mounted and nothing prebuilt)")
    return oracle


@pytest.mark.parametrize("clustered", [False, True])
@pytest.mark.parametrize("d,nlist,nq,nprobe", CASES)
def test_oracle_coarse_matches_the_references_own_code(ref, d, nlist, nq, nprobe, clustered):
    cent, xq = _data(11, d, nlist, nq, clustered)
    Dr, Ir = ref.ref_coarse(xq, cent, nprobe)
    Do, Io = ref.C.coarse(xq, cent, nprobe)
    assert np.all(np.diff(Dr, axis=1) >= 0), "reference rows ascending"
    _util.assert_same_modulo_near_ties(Do, Io, Dr, Ir, RTOL, f"oracle vs hnswlib d={d} nlist={nlist}")
    # away from ties the two are the same ranking
    assert (Io == Ir).mean() > 0.999


def test_exact_ties_are_the_only_freedom(ref):
    """Duplicate centroids: both sides return the same distance profile; which twin comes first is the tie rule
    (oracle: lower id; hnswlib: heap order) -- the only place the ids may differ."""
    cent, xq = _data(5, 64, 128, 40, False)
    cent[64:] = cent[:64]
    Dr, Ir = ref.ref_coarse(xq, cent, 10)
    Do, Io = ref.C.coarse(xq, cent, 10)
    np.testing.assert_allclose(Do, Dr, rtol=RTOL)
    assert np.array_equal(np.sort(Io % 64, axis=1), np.sort(Ir % 64, axis=1))
    assert np.all(Io[:, 0::2] < 64), "oracle tie rule: the lower id of a twin pair first"


@pytest.mark.gpu
@pytest.mark.parametrize("clustered", [False, True])
@pytest.mark.parametrize("d,nlist,nq,nprobe", CASES[:4])
def test_gpu_coarse_matches_the_references_own_code(ref, d, nlist, nq, nprobe, clustered):
    """The CUDA coarse quantizer (tcgen05 pre-filter + exact rescoring, or the small-batch kernels) against the
    reference's hnswlib selection on the same inputs."""
    import b200ivfpq as faiss
    cent, xq = _data(11, d, nlist, nq, clustered)
    q = faiss.IndexFlatL2(d)
    q.add(cent)
    D, I = q.search(xq, nprobe)
    Dr, Ir = ref.ref_coarse(xq, cent, nprobe)
    _util.assert_same_modulo_near_ties(np.asarray(D), np.asarray(I), Dr, Ir, RTOL, f"GPU vs hnswlib d={d} nlist={nlist}")
