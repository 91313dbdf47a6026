"""Parity of the CUDA path (through the C-ABI) with the oracle, on the same codebooks and inputs.

Bar (BASELINE.json north_star): probed lists identical, returned ids identical, distances bit-exact
(tighter than the 1e-5 relative the north_star allows: both sides follow the same fp32 operation order).
"""
import os

import numpy as np
import pytest

import _util

pytestmark = pytest.mark.gpu

RTOL = 1e-5   # the tolerance BASELINE.json states; the tests below assert bit-exactness, which implies it


def _load(a, variant=None):
    import b200ivfpq as faiss
    old = os.environ.get("B200_IVFPQ_SCAN")
    if variant:
        os.environ["B200_IVFPQ_SCAN"] = variant
    try:
        index = faiss.IndexIVFPQ(faiss.IndexFlatL2(a["d"]), a["d"], a["nlist"], a["M"], 8)
        index.set_codebooks(a["coarse"], a["pq"])
        index.set_lists(a["offsets"], a["codes"], a["ids"])
    finally:
        if variant:
            if old is None:
                del os.environ["B200_IVFPQ_SCAN"]
            else:
                os.environ["B200_IVFPQ_SCAN"] = old
    return index


CASES = [
    # d, nlist, M, n, nq, nprobe, k, used_lists
    (128, 64, 16, 20000, 64, 8, 10, None),      # SIFT shape, skewed-lane kernel
    (96, 32, 16, 9000, 33, 6, 100, None),       # Deep1B shape (dsub 6), k = 100
    (128, 16, 16, 5000, 7, 16, 10, 11),         # empty + ragged lists, nprobe == nlist
    (64, 40, 16, 300, 20, 5, 10, None),         # lists shorter than one warp block
    (128, 50, 32, 8000, 40, 7, 10, None),       # M = 32
    (768, 12, 64, 3000, 9, 4, 10, None),        # RALM shape, M = 64
    (32, 24, 8, 3000, 16, 6, 10, 20),           # M = 8 (IVFPQ_random_dataset.py uses m = 8)
    (20, 10, 5, 900, 11, 3, 5, None),           # M = 5: byte-wise code loads
    (24, 9, 12, 700, 5, 9, 1, None),            # k = 1
]


@pytest.mark.parametrize("d,nlist,M,n,nq,nprobe,k,used", CASES)
def test_search_bit_exact(oracle, d, nlist, M, n, nq, nprobe, k, used):
    a = _util.make_index_arrays(oracle, 100 + d + M, d, nlist, M, n, used_lists=used)
    xq = _util.make_queries(7, a, nq)
    Dr, Ir, pdis, pid = oracle.C.search(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], nprobe, k,
                                        return_probes=True)
    index = _load(a)
    index.nprobe = nprobe
    # a1: probed lists identical, coarse distances bit-exact
    cd, ci = index.quantizer.search(xq, nprobe)
    _util.assert_bit_equal(ci, pid, "probed list ids")
    _util.assert_bit_equal(cd, pdis, "coarse distances")
    # numpy in / numpy out: the C-ABI host entry point
    D, I = index.search(xq, k)
    _util.assert_bit_equal(D, Dr, "D (host path)")
    _util.assert_bit_equal(I, Ir, "I (host path)")
    np.testing.assert_allclose(D, Dr, rtol=RTOL)
    # torch in / torch out: device pointers on the current stream
    import torch
    Dt, It = index.search(torch.from_numpy(xq).cuda(), k)
    _util.assert_bit_equal(Dt.cpu().numpy(), Dr, "D (device path)")
    _util.assert_bit_equal(It.cpu().numpy(), Ir, "I (device path)")


@pytest.mark.parametrize("variant", ["generic", "skew", "duo", "quad"])
def test_both_scan_kernels_agree_with_oracle(oracle, variant):
    a = _util.make_index_arrays(oracle, 5, 128, 48, 16, 30000)
    xq = _util.make_queries(8, a, 50)
    Dr, Ir = oracle.C.search(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], 12, 20)
    index = _load(a, variant)
    index.nprobe = 12
    D, I = index.search(xq, 20)
    _util.assert_bit_equal(D, Dr, f"D ({variant})")
    _util.assert_bit_equal(I, Ir, f"I ({variant})")


@pytest.mark.parametrize("variant", ["generic", "skew", "duo", "quad"])
def test_ties_follow_scan_order(oracle, variant):
    """Many duplicate codes -> equal distances; (distance, probe rank, offset) order must match the oracle."""
    rng = np.random.default_rng(3)
    a = _util.make_index_arrays(oracle, 6, 64, 8, 16, 4000)
    few = a["codes"][rng.integers(0, 40, size=a["codes"].shape[0])]     # only 40 distinct codes
    a["codes"] = np.ascontiguousarray(few)
    xq = _util.make_queries(2, a, 30)
    Dr, Ir = oracle.C.search(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], 8, 50)
    assert any(len(set(_util.bits(r).tolist())) < 50 for r in Dr), "test needs real ties"
    index = _load(a, variant)
    index.nprobe = 8
    D, I = index.search(xq, 50)
    _util.assert_bit_equal(D, Dr, "D")
    _util.assert_bit_equal(I, Ir, "I")


@pytest.mark.parametrize("d,nlist,n,nq,nprobe,k,used", [
    (128, 24, 30000, 101, 5, 10, None),     # odd numbers of queries per list: groups with a missing second query
    (96, 16, 12000, 64, 16, 100, 13),       # dsub 6, k = 100, empty lists, every query probes every list
    (64, 8, 3000, 1, 8, 10, None),          # one query: every group is a single
    (256, 12, 5000, 37, 3, 7, None),        # dsub 16
    (80, 12, 5000, 40, 4, 10, None),        # dsub 5: generic LUT build
])
def test_two_query_scan_kernel(oracle, d, nlist, n, nq, nprobe, k, used):
    """scan_duo.cuh: two queries of the same list per work item; results must not depend on the grouping."""
    a = _util.make_index_arrays(oracle, 40 + d, d, nlist, 16, n, used_lists=used)
    xq = _util.make_queries(11, a, nq)
    Dr, Ir = oracle.C.search(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], nprobe, k)
    index = _load(a, "duo")
    index.nprobe = nprobe
    for _ in range(2):          # the pairing of queries depends on atomics' order; the results must not
        D, I = index.search(xq, k)
        _util.assert_bit_equal(D, Dr, "D (duo)")
        _util.assert_bit_equal(I, Ir, "I (duo)")


@pytest.mark.parametrize("d,nlist,n,nq,nprobe,k,used,M", [
    (128, 24, 60000, 103, 5, 10, None, 16),     # group sizes 1..4, several tiles per list
    (96, 16, 12000, 64, 16, 100, 13, 16),       # dsub 6, k = 100, empty lists, every query probes every list
    (64, 8, 3000, 1, 8, 10, None, 16),          # one query: every group is a single
    (256, 12, 5000, 37, 3, 7, None, 16),        # dsub 16: generic LUT build
    (128, 4, 9000, 200, 4, 1, None, 16),        # k = 1, 200 queries on every list
    (128, 24, 60000, 103, 5, 10, None, 32),     # M = 32 (C5 shape, dsub 4): two tables, 10-bit entries
    (256, 16, 12000, 64, 16, 100, 13, 32),      # M = 32, dsub 8, k = 100, empty lists
    (192, 6, 2500, 9, 6, 10, None, 32),         # M = 32, dsub 6: generic LUT build
])
def test_four_query_filter_scan_kernel(oracle, d, nlist, n, nq, nprobe, k, used, M):
    """scan_quad.cuh: integer lower-bound filter + exact evaluation of the survivors.  Must return exactly the oracle's
    results, ties included (many duplicate codes make sure survivors at the threshold are not lost)."""
    a = _util.make_index_arrays(oracle, 90 + d, d, nlist, M, n, used_lists=used)
    if d == 128 and nlist == 4:
        rng = np.random.default_rng(5)
        a["codes"] = np.ascontiguousarray(a["codes"][rng.integers(0, 500, size=a["codes"].shape[0])])   # heavy ties
    xq = _util.make_queries(17, a, nq)
    Dr, Ir = oracle.C.search(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], nprobe, k)
    index = _load(a, "quad")
    index.nprobe = nprobe
    for _ in range(2):
        D, I = index.search(xq, k)
        _util.assert_bit_equal(D, Dr, "D (quad)")
        _util.assert_bit_equal(I, Ir, "I (quad)")


@pytest.mark.parametrize("d,nlist,n,nq,nprobe,k,used", [
    (128, 24, 40000, 101, 5, 10, None),     # C5 shape (dsub 4), odd group sizes, several tiles per list
    (256, 16, 9000, 64, 16, 100, 13),       # dsub 8, k = 100, empty lists
    (512, 8, 3000, 1, 8, 10, None),         # dsub 16 (RALM-S shape), one query: every group is a single
    (768, 12, 5000, 37, 3, 7, None),        # dsub 24: generic LUT build
    (64, 6, 700, 20, 6, 10, None),          # dsub 2, lists shorter than one 512-code iteration
])
def test_two_query_scan_kernel_m32(oracle, d, nlist, n, nq, nprobe, k, used):
    """scan_duo32.cuh: M = 32, two alternating look-up tables, two queries per work item."""
    a = _util.make_index_arrays(oracle, 70 + d, d, nlist, 32, n, used_lists=used)
    xq = _util.make_queries(13, a, nq)
    Dr, Ir = oracle.C.search(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], nprobe, k)
    index = _load(a, "duo")
    index.nprobe = nprobe
    for _ in range(2):
        D, I = index.search(xq, k)
        _util.assert_bit_equal(D, Dr, "D (duo32)")
        _util.assert_bit_equal(I, Ir, "I (duo32)")


@pytest.mark.parametrize("variant,M", [("auto", 16), ("skew", 16), ("auto", 32), ("generic", 8)])
def test_large_k(oracle, variant, M):
    """k = 1000 (Faiss-GPU allows up to 2048): the candidate queues fold through the shared-memory bitonic path, most
    lists return fewer than k results, and unfilled slots come back as FLT_MAX / -1."""
    d = 64 if M != 32 else 128
    a = _util.make_index_arrays(oracle, 33 + M, d, 16, M, 30000)
    xq = _util.make_queries(19, a, 48)
    for nprobe in (8, 1):
        Dr, Ir = oracle.C.search(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], nprobe, 1000)
        index = _load(a, None if variant == "auto" else variant)
        index.nprobe = nprobe
        D, I = index.search(xq, 1000)
        _util.assert_bit_equal(D, Dr, f"D (k = 1000, {variant}, nprobe {nprobe})")
        _util.assert_bit_equal(I, Ir, f"I (k = 1000, {variant}, nprobe {nprobe})")


def test_search_preassigned(oracle):
    import b200ivfpq as faiss
    a = _util.make_index_arrays(oracle, 9, 128, 20, 16, 6000)
    xq = _util.make_queries(4, a, 12)
    rng = np.random.default_rng(0)
    pid = rng.integers(-1, 20, size=(12, 5)).astype(np.int64)
    pid[3] = -1
    Dr, Ir = oracle.C.search_preassigned(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], pid, 10)
    index = _load(a)
    D, I = faiss.search_preassigned(index, xq, 10, pid)
    _util.assert_bit_equal(D, Dr)
    _util.assert_bit_equal(I, Ir)
    assert (I[3] == -1).all() and (D[3] == oracle.FLT_MAX).all()


def test_unfilled_and_empty_index(oracle):
    a = _util.make_index_arrays(oracle, 4, 16, 6, 4, 9)
    xq = _util.make_queries(1, a, 3)
    Dr, Ir = oracle.C.search(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], 1, 16)
    index = _load(a)
    D, I = index.search(xq, 16)
    _util.assert_bit_equal(D, Dr)
    _util.assert_bit_equal(I, Ir)
    a["offsets"] = np.zeros(7, np.int64)
    a["codes"] = np.zeros((0, 4), np.uint8)
    a["ids"] = np.zeros(0, np.int64)
    index = _load(a)
    D, I = index.search(xq, 4)
    assert (I == -1).all() and (D == oracle.FLT_MAX).all()


def test_golden_search_small(oracle):
    z = np.load(os.path.join(_util.GOLDEN, "search_small.npz"))
    a = {"d": 32, "nlist": 24, "M": 8, "coarse": z["coarse"], "pq": z["pq"], "offsets": z["offsets"],
         "codes": z["codes"], "ids": z["ids"]}
    index = _load(a)
    index.nprobe = int(z["nprobe"])
    D, I = index.search(z["xq"], int(z["k"]))
    _util.assert_bit_equal(D, z["D"])
    _util.assert_bit_equal(I, z["I"])


def test_lut_kat_through_the_kernels():
    """The reference's literal LUT known-answer test (host.cpp:44-109) pushed through K2+K3: list m holds the 256
    codes that are zero except byte m = k, so the returned distances are T[m][k] + sum_{m' != m} T[m'][0].
    All values are integers < 2^24, hence exact in fp32 in any order."""
    z = np.load(os.path.join(_util.GOLDEN, "lut_kat_d128_m32.npz"))
    T = z["lut"].astype(np.int64)                     # (32, 256)
    M, d = 32, 128
    nlist = M
    coarse = np.tile(z["centroid"][None, :], (nlist, 1)).astype(np.float32)
    codes = np.zeros((M * 256, M), np.uint8)
    for m in range(M):
        codes[m * 256:(m + 1) * 256, m] = np.arange(256)
    offsets = np.arange(0, M * 256 + 1, 256, dtype=np.int64)
    a = {"d": d, "nlist": nlist, "M": M, "coarse": coarse, "pq": z["pq"], "offsets": offsets, "codes": codes,
         "ids": np.arange(M * 256, dtype=np.int64)}
    index = _load(a)
    xq = z["query"][None, :].astype(np.float32)
    base = T[:, 0].sum()
    for m in (0, 1, 13, 31):
        D, I = index.search_preassigned(xq, 256, np.array([[m]], np.int64))
        expect = base - T[m, 0] + T[m]                # distance of code k in list m
        got = np.empty(256, np.int64)
        got[I[0] - m * 256] = D[0].astype(np.int64)
        assert np.array_equal(got, expect), f"list {m}"
        assert (np.diff(D[0]) >= 0).all()


def test_assign_encode_bit_exact(oracle):
    import b200ivfpq as faiss
    a = _util.make_index_arrays(oracle, 31, 96, 40, 16, 5000)
    index = faiss.IndexIVFPQ(faiss.IndexFlatL2(96), 96, 40, 16, 8)
    index.set_codebooks(a["coarse"], a["pq"])
    index.add_with_ids(a["x"], np.arange(5000, dtype=np.int64) * 7 + 3)
    out = index.to_arrays()
    order = np.argsort(a["list_no"], kind="stable")
    codes_ref = oracle.C.encode(a["x"], a["coarse"], a["list_no"], a["pq"])[order]
    _util.assert_bit_equal(out["offsets"], a["offsets"], "offsets (assignment)")
    _util.assert_bit_equal(out["codes"], codes_ref, "codes")
    _util.assert_bit_equal(out["ids"], (np.arange(5000, dtype=np.int64) * 7 + 3)[order], "ids")


def test_c1_shape_properties(oracle):
    """BASELINE config 1 shape at full size (1M x 128, IVF1024, PQ16, nprobe 16, k 10): random codes in the CSR
    layout; full oracle comparison on a query sample + size-independent properties on all 10k queries."""
    import torch
    rng = np.random.default_rng(1234)
    d, nlist, M, n, nq, nprobe, k = 128, 1024, 16, 1_000_000, 10_000, 16, 10
    coarse = rng.random((nlist, d), dtype=np.float32)
    pq = (rng.standard_normal((M, 256, d // M)) * 0.1).astype(np.float32)
    sizes = rng.multinomial(n, rng.dirichlet(np.full(nlist, 2.0)))
    offsets = np.zeros(nlist + 1, np.int64)
    offsets[1:] = np.cumsum(sizes)
    codes = rng.integers(0, 256, size=(n, M), dtype=np.uint8)
    ids = rng.permutation(n).astype(np.int64)
    a = {"d": d, "nlist": nlist, "M": M, "coarse": coarse, "pq": pq, "offsets": offsets, "codes": codes, "ids": ids}
    xq = rng.random((nq, d), dtype=np.float32)
    index = _load(a)
    index.nprobe = nprobe
    D, I = index.search(xq, k)
    # properties on the whole batch
    assert (np.diff(D, axis=1) >= 0).all(), "rows ascending"
    assert (I >= 0).all() and (I < n).all()
    assert all(len(set(r.tolist())) == k for r in I[::97]), "ids distinct within a row"
    D2, I2 = index.search(xq, k)
    _util.assert_bit_equal(D2, D, "idempotent D")
    _util.assert_bit_equal(I2, I, "idempotent I")
    # batch-size independence: any sub-batch gives the same rows
    D3, I3 = index.search(xq[123:124], k)
    _util.assert_bit_equal(D3[0], D[123])
    _util.assert_bit_equal(I3[0], I[123])
    # oracle on a sample
    sel = np.arange(0, nq, 50)
    Dr, Ir = oracle.C.search(xq[sel], coarse, pq, offsets, codes, ids, nprobe, k)
    _util.assert_bit_equal(D[sel], Dr, "D vs oracle")
    _util.assert_bit_equal(I[sel], Ir, "I vs oracle")
    st = index.last_scan_stats()
    assert st["bytes"] == st["codes"] * M


@pytest.mark.parametrize("M,d", [(16, 128), (8, 32)])
def test_small_batches_graph_replay_and_segments(oracle, M, d):
    """Batch-1 latency path: list segmentation (each pair split over several CTAs) and CUDA-graph replay of the
    host-buffer search.  Repeated calls with DIFFERENT queries must keep matching the oracle, and an index update
    must invalidate the recorded graph."""
    a = _util.make_index_arrays(oracle, 77, d, 40, M, 30000)
    index = _load(a)
    index.nprobe = 7
    for nq in (1, 3, 17):
        for rep in range(4):                              # call 1-2: plain, call 3+: graph replay
            xq = _util.make_queries(1000 * nq + rep, a, nq)
            Dr, Ir = oracle.C.search(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], 7, 10)
            D, I = index.search(xq, 10)
            _util.assert_bit_equal(D, Dr, f"D nq={nq} rep={rep}")
            _util.assert_bit_equal(I, Ir, f"I nq={nq} rep={rep}")
    # mutate the index: new lists -> the graph recorded for (nq=1, k=10, nprobe=7) must not be replayed
    b = _util.make_index_arrays(oracle, 78, d, 40, M, 9000)
    index.set_codebooks(b["coarse"], b["pq"])
    index.set_lists(b["offsets"], b["codes"], b["ids"])
    xq = _util.make_queries(5, b, 1)
    Dr, Ir = oracle.C.search(xq, b["coarse"], b["pq"], b["offsets"], b["codes"], b["ids"], 7, 10)
    for rep in range(3):
        D, I = index.search(xq, 10)
        _util.assert_bit_equal(D, Dr, "D after update")
        _util.assert_bit_equal(I, Ir, "I after update")


@pytest.mark.parametrize("variant,nseg", [("skew", 3), ("skew", 16), ("generic", 5)])
def test_forced_list_segmentation(oracle, variant, nseg):
    a = _util.make_index_arrays(oracle, 21, 128, 24, 16, 26000, used_lists=20)
    xq = _util.make_queries(4, a, 40)
    Dr, Ir = oracle.C.search(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], 9, 25)
    os.environ["B200_IVFPQ_NSEG"] = str(nseg)
    try:
        index = _load(a, variant)
    finally:
        del os.environ["B200_IVFPQ_NSEG"]
    index.nprobe = 9
    import torch
    D, I = index.search(torch.from_numpy(xq).cuda(), 25)
    _util.assert_bit_equal(D.cpu().numpy(), Dr, "D")
    _util.assert_bit_equal(I.cpu().numpy(), Ir, "I")
