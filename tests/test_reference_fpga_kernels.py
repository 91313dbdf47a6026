"""Rows a2-a4 against CODE OF THE REFERENCE run here: the HLS kernels LUT_construction.hpp (LUT_construction_wrapper)
and ADC.hpp (PQ_lookup_computation) of retrieval_accelerator/entire_accelerator_final_{SIFT,Deep}_M{16,32}/src, compiled
with g++ as a C simulation (oracle/ref_fpga_shim.cpp + our stand-ins for Xilinx's ap_int.h / hls_stream.h) into
oracle/_ref/.  Their arithmetic -- residual, (r - p)^2 summed sequentially, table entries added in ascending m -- is
the contract of the oracle and of the CUDA kernels, so everything here is BIT-exact."""
import os

import numpy as np
import pytest

import _util

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VARIANTS = ["SIFT_M16", "SIFT_M32", "Deep_M16", "Deep_M32"]     # (D, M) = (128,16) (128,32) (96,16) (96,32)


@pytest.fixture(scope="module")
def ref(oracle):
    if oracle.build_ref() is None:
        pytest.skip("oracle/_ref not available (reference not mounted and nothing prebuilt)")
    for v in VARIANTS:
        if not os.path.exists(os.path.join(os.path.dirname(oracle.__file__), "_ref", f"libref_fpga_{v}.so")):
            pytest.skip(f"oracle/_ref/libref_fpga_{v}.so missing")
    return oracle


def _case(oracle, variant, seed, nq=3, nprobe=5, nlist=24, max_list=60, scale=1.0):
    D, M = oracle.FPGA_VARIANTS[variant]
    rng = np.random.default_rng(seed)
    pq = (rng.standard_normal((M, 256, D // M)) * 0.3 * scale).astype(np.float32)
    cent = (rng.random((nlist, D), dtype=np.float32) * np.float32(scale)).astype(np.float32)
    xq = (rng.random((nq, D), dtype=np.float32) * np.float32(scale)).astype(np.float32)
    sizes = rng.integers(0, max_list, nlist)
    sizes[rng.integers(0, nlist, 3)] = 0                       # some empty cells
    offsets = np.zeros(nlist + 1, np.int64)
    offsets[1:] = np.cumsum(sizes)
    codes = rng.integers(0, 256, (int(offsets[-1]), M), dtype=np.uint8)
    ids = rng.permutation(int(offsets[-1])).astype(np.int64) + 1000
    _, probes = oracle.C.coarse(xq, cent, nprobe)
    return dict(D=D, M=M, pq=pq, cent=cent, xq=xq, offsets=offsets, codes=codes, ids=ids, probes=probes)


def _run_reference_kernels(oracle, variant, c):
    """Feed the reference's kernels what its host would: per query the probed cells' centroids and codes."""
    probes, off = c["probes"], c["offsets"]
    centers = c["cent"][probes]
    nscan = (off[probes + 1] - off[probes]).astype(np.int32)
    rows = np.concatenate([np.arange(off[l], off[l + 1]) for l in probes.ravel()] + [np.empty(0, np.int64)])
    rows = rows.astype(np.int64)
    lut, dist = oracle.ref_fpga_lut_adc(variant, c["pq"], c["xq"], centers, nscan, c["codes"][rows])
    return lut, dist, nscan, rows


@pytest.mark.parametrize("scale", [1.0, 255.0])
@pytest.mark.parametrize("variant", VARIANTS)
def test_oracle_lut_and_adc_are_bit_identical_to_the_references_hls_kernels(ref, variant, scale):
    c = _case(ref, variant, seed=7, scale=scale)
    lut, dist, nscan, rows = _run_reference_kernels(ref, variant, c)
    pos = 0
    for q in range(c["xq"].shape[0]):
        for p, l in enumerate(c["probes"][q]):
            T = ref.C.lut(c["xq"][q], c["cent"][l], c["pq"])                       # (M, 256)
            _util.assert_bit_equal(T, lut[q, p], f"{variant} LUT q{q} probe{p}")
            n = int(nscan[q, p])
            if n:
                d = ref.C.adc(T, c["codes"][rows[pos:pos + n]])
                _util.assert_bit_equal(d, dist[pos:pos + n], f"{variant} ADC q{q} probe{p}")
            pos += n
    assert pos == dist.shape[0]


def test_the_references_own_kat_input_through_its_own_kernel(ref):
    """LUT_construction_PE_D128_M32/src/host.cpp:44-109 (tests/golden/lut_kat_d128_m32.npz holds its literal inputs and the
    exact expected table): the full-accelerator build of the same kernel reproduces it bit for bit."""
    g = np.load(os.path.join(GOLDEN, "lut_kat_d128_m32.npz"))
    lut, _ = ref.ref_fpga_lut_adc("SIFT_M32", g["pq"], g["query"][None], g["centroid"][None, None],
                                  np.zeros((1, 1), np.int32), np.empty((0, 32), np.uint8))
    _util.assert_bit_equal(lut[0, 0], g["lut"], "reference kernel vs its own known-answer table")


def _check_search_against_reference_kernels(ref, variant, search_preassigned):
    """End to end: with k = everything scanned, search_preassigned returns every candidate's distance; the sorted
    distances must be the reference kernels' ADC outputs bit for bit, and every id must carry its own distance."""
    c = _case(ref, variant, seed=21, nq=4, nprobe=6, nlist=20, max_list=40)
    _, dist, nscan, rows = _run_reference_kernels(ref, variant, c)
    k = int(nscan.sum(axis=1).max())
    assert 1 <= k <= 1000
    D, I = search_preassigned(c, k)
    pos = 0
    for q in range(c["xq"].shape[0]):
        n = int(nscan[q].sum())
        want = np.sort(dist[pos:pos + n])
        _util.assert_bit_equal(np.asarray(D[q, :n], np.float32), want, f"{variant} q{q}: all scanned distances")
        assert np.all(np.asarray(I[q, n:]) == -1)
        by_id = dict(zip(c["ids"][rows[pos:pos + n]].tolist(), dist[pos:pos + n].tolist()))
        for dq, iq in zip(np.asarray(D[q, :n]).tolist(), np.asarray(I[q, :n]).tolist()):
            assert by_id[iq] == dq, f"{variant} q{q}: id {iq} carries another entry's distance"
        pos += n


@pytest.mark.parametrize("variant", VARIANTS)
def test_oracle_search_returns_the_reference_kernels_distances(ref, variant):
    def run(c, k):
        return ref.C.search_preassigned(c["xq"], c["cent"], c["pq"], c["offsets"], c["codes"], c["ids"], c["probes"], k)
    _check_search_against_reference_kernels(ref, variant, run)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", VARIANTS)
def test_cuda_search_returns_the_reference_kernels_distances(ref, variant):
    """The product path (K2+K3+K4 on the GPU) against the reference's HLS kernels, bit for bit."""
    import b200ivfpq as faiss

    def run(c, k):
        index = faiss.IndexIVFPQ(faiss.IndexFlatL2(c["D"]), c["D"], c["cent"].shape[0], c["M"], 8)
        index.set_codebooks(c["cent"], c["pq"])
        index.set_lists(c["offsets"], c["codes"], c["ids"])
        return index.search_preassigned(c["xq"], k, c["probes"])
    _check_search_against_reference_kernels(ref, variant, run)


# ---- row a5: the selection rule, against the reference's own queue ------------------------------------------------
def _tie_case(oracle, variant, seed, n=400, distinct=6, nq=12, nlist=8):
    """ONE populated cell (the others are empty) whose entries repeat a handful of distinct codes: every query sees long
    runs of EXACTLY equal distances, so which entries make the top-k is decided by the tie rule alone."""
    D, M = oracle.FPGA_VARIANTS[variant]
    rng = np.random.default_rng(seed)
    pq = (rng.standard_normal((M, 256, D // M)) * 0.3).astype(np.float32)
    cent = rng.random((nlist, D), dtype=np.float32)
    xq = rng.random((nq, D), dtype=np.float32)
    rows = rng.integers(0, 256, (distinct, M), dtype=np.uint8)
    codes = rows[rng.integers(0, distinct, n)]
    offsets = np.full(nlist + 1, n, np.int64)
    offsets[0] = 0                                               # everything lives in cell 0
    return dict(D=D, M=M, pq=pq, cent=cent, xq=xq, offsets=offsets, codes=codes,
                ids=np.arange(n, dtype=np.int64), probes=np.zeros((nq, 1), np.int64))


def _check_selection_against_reference_queue(ref, variant, k, search_preassigned):
    """Reference chain, all its own code: HLS LUT -> HLS ADC -> systolic queue of length k (strict `<`,
    priority_queue_L1.hpp:65-75) fed in scan order.  The search must return the same k entries: with ids = scan
    positions, the same ids in the same (distance, scan order) ranking, the same distance bits."""
    c = _tie_case(ref, variant, seed=3 + k)
    n = c["codes"].shape[0]
    nq = c["xq"].shape[0]
    _, dist = ref.ref_fpga_lut_adc(variant, c["pq"], c["xq"], c["cent"][c["probes"]], np.full((nq, 1), n, np.int32),
                                   np.tile(c["codes"], (nq, 1)))
    D, I = search_preassigned(c, k)
    for q in range(nq):
        dq = dist[q * n:(q + 1) * n]
        assert np.unique(dq).shape[0] <= 6, "the case is built from at most six distinct distances"
        off, od = ref.ref_fpga_queue(dq, k, variant)
        assert off.shape[0] == min(k, n)
        assert np.array_equal(np.asarray(I[q, :off.shape[0]]), off), f"{variant} k={k} q{q}: selected entries differ"
        _util.assert_bit_equal(np.asarray(D[q, :off.shape[0]], np.float32), od, f"{variant} k={k} q{q}: distances")


@pytest.mark.parametrize("k", [1, 10, 100])
@pytest.mark.parametrize("variant", ["SIFT_M16", "Deep_M16", "SIFT_M32"])
def test_oracle_tie_rule_is_the_references_queue(ref, variant, k):
    def run(c, kk):
        return ref.C.search_preassigned(c["xq"], c["cent"], c["pq"], c["offsets"], c["codes"], c["ids"], c["probes"], kk)
    _check_selection_against_reference_queue(ref, variant, k, run)


def test_reference_queue_is_exact_under_heavy_ties(ref):
    """The systolic queue by itself against the rule it is taken to implement: the k smallest under the total order
    (distance, scan order) -- among equal distances the earlier-scanned entry stays."""
    rng = np.random.default_rng(0)
    for k in (1, 10, 100):
        for n, levels in [(5, 3), (50, 8), (500, 20), (3000, 50), (3000, 100000)]:
            d = (rng.integers(0, levels, n) / levels).astype(np.float32)
            off, od = ref.ref_fpga_queue(d, k)
            want = np.lexsort((np.arange(n), d))[:k]
            assert np.array_equal(off, want), (k, n, levels)
            _util.assert_bit_equal(od, d[want], "queue distances")


@pytest.mark.gpu
@pytest.mark.parametrize("k", [1, 10, 100])
@pytest.mark.parametrize("variant", ["SIFT_M16", "Deep_M16", "SIFT_M32"])
def test_cuda_tie_rule_is_the_references_queue(ref, variant, k):
    import b200ivfpq as faiss

    def run(c, kk):
        index = faiss.IndexIVFPQ(faiss.IndexFlatL2(c["D"]), c["D"], c["cent"].shape[0], c["M"], 8)
        index.set_codebooks(c["cent"], c["pq"])
        index.set_lists(c["offsets"], c["codes"], c["ids"])
        return index.search_preassigned(c["xq"], kk, c["probes"])
    _check_selection_against_reference_queue(ref, variant, k, run)


# ---- randomised shapes ------------------------------------------------------------------------------------------------
def test_random_shapes_against_the_reference_kernels(ref):
    """Property test (hypothesis): for random batch sizes, probe counts, cell sizes (empty cells included) and value
    scales, the oracle's ADC distances are bit-identical to the reference HLS kernels' outputs."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None, derandomize=True)
    @given(variant=st.sampled_from(VARIANTS), seed=st.integers(0, 2 ** 31 - 1), nq=st.integers(1, 3),
           nprobe=st.integers(1, 6), max_list=st.integers(1, 70), scale=st.sampled_from([1e-3, 1.0, 255.0, 1e4]))
    def run(variant, seed, nq, nprobe, max_list, scale):
        c = _case(ref, variant, seed=seed, nq=nq, nprobe=nprobe, nlist=12, max_list=max_list, scale=scale)
        lut, dist, nscan, rows = _run_reference_kernels(ref, variant, c)
        pos = 0
        for q in range(nq):
            for p, l in enumerate(c["probes"][q]):
                n = int(nscan[q, p])
                if n:
                    T = ref.C.lut(c["xq"][q], c["cent"][l], c["pq"])
                    _util.assert_bit_equal(T, lut[q, p], "LUT")
                    _util.assert_bit_equal(ref.C.adc(T, c["codes"][rows[pos:pos + n]]), dist[pos:pos + n], "ADC")
                pos += n

    run()
