"""Parity of the per-query-table filter scan (csrc/scan_qlut.cuh) with the oracle.

The kernel filters codes with an integer lower bound built from a per-QUERY table and a per-VECTOR 16-bit term, and
evaluates the survivors exactly from the codebook.  It must return exactly the oracle's results -- distances bit for
bit, ids, ties in scan order -- whatever the data looks like: the cases below include SIFT-scale values, data far from
the origin (where the decomposition cancels badly), the reference's own toy generator with its 0..n/1000 ramp on
dimension 0 (IVFPQ_random_dataset.py:6-13), duplicate codes, empty lists, k = 1 .. 500 and M = 16 / 32 / 64.
"""
import ctypes
import os

import numpy as np
import pytest

import _util

pytestmark = pytest.mark.gpu


def _load(a, variant="qlut", stats=False, stream=1, env=None):
    """stream = 1: the streaming pipeline (csrc/scan_stream.cuh: bootstrap thresholds, filter, exact evaluation, select,
    with the in-kernel path as its overflow fallback); stream = 0: the in-kernel top-k path (scan_qlut_kernel) alone."""
    import b200ivfpq as faiss
    names = ("B200_IVFPQ_SCAN", "B200_IVFPQ_QL_STATS", "B200_IVFPQ_STREAM", "B200_IVFPQ_STREAM_MINREC",
             "B200_IVFPQ_STREAM_RATE", "B200_IVFPQ_STREAM_TWO")
    old = {k: os.environ.get(k) for k in names}
    os.environ["B200_IVFPQ_SCAN"] = variant
    os.environ["B200_IVFPQ_STREAM"] = str(stream)
    if stats:
        os.environ["B200_IVFPQ_QL_STATS"] = "1"
    for k_, v_ in (env or {}).items():
        os.environ[k_] = v_
    try:
        index = faiss.IndexIVFPQ(faiss.IndexFlatL2(a["d"]), a["d"], a["nlist"], a["M"], 8)
        index.set_codebooks(a["coarse"], a["pq"])
        index.set_lists(a["offsets"], a["codes"], a["ids"])
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return index


def _check(oracle, a, xq, nprobe, k, what, repeats=2):
    Dr, Ir = oracle.C.search(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], nprobe, k)
    for stream in (1, 0):
        index = _load(a, stream=stream)
        index.nprobe = nprobe
        for _ in range(repeats):      # the grouping of queries depends on atomics' order; the results must not
            D, I = index.search(xq, k)
            _util.assert_bit_equal(D, Dr, f"D ({what}, stream={stream})")
            _util.assert_bit_equal(I, Ir, f"I ({what}, stream={stream})")
    return index


@pytest.mark.parametrize("d,nlist,n,nq,nprobe,k,used,M", [
    (128, 24, 60000, 103, 5, 10, None, 16),     # group sizes 1..4, several tiles per list
    (96, 16, 12000, 64, 16, 100, 13, 16),       # dsub 6, k = 100, empty lists, every query probes every list
    (64, 8, 3000, 1, 8, 10, None, 16),          # one query: every group is a single
    (256, 12, 5000, 37, 3, 7, None, 16),        # dsub 16
    (80, 12, 5000, 40, 4, 10, None, 16),        # dsub 5
    (128, 4, 9000, 200, 4, 1, None, 16),        # k = 1, 200 queries on every list
    (128, 24, 60000, 103, 5, 10, None, 32),     # M = 32 (C5 shape, dsub 4): two chunk tables, sums widened per chunk
    (256, 16, 12000, 64, 16, 100, 13, 32),      # M = 32, dsub 8, k = 100, empty lists
    (192, 6, 2500, 9, 6, 10, None, 32),         # M = 32, dsub 6
    (768, 12, 9000, 40, 4, 10, None, 64),       # M = 64 (C4 / RALM shape, dsub 12): four chunk tables
    (128, 6, 2000, 9, 6, 20, 5, 64),            # M = 64, dsub 2, an empty list
    (128, 10, 30000, 64, 10, 500, None, 16),    # k = 500: shared-memory bitonic folds
])
def test_qlut_scan_bit_exact(oracle, d, nlist, n, nq, nprobe, k, used, M):
    a = _util.make_index_arrays(oracle, 190 + d + M, d, nlist, M, n, used_lists=used)
    if d == 128 and nlist == 4:
        rng = np.random.default_rng(5)
        a["codes"] = np.ascontiguousarray(a["codes"][rng.integers(0, 500, size=a["codes"].shape[0])])   # heavy ties
    xq = _util.make_queries(17, a, nq)
    _check(oracle, a, xq, nprobe, k, f"qlut M={M}")


def _scaled(oracle, seed, d, nlist, M, n, nq, scale, shift):
    """Clustered data times `scale` plus `shift` on every coordinate; codebooks trained by a few Lloyd steps so that the
    codes are meaningful at that scale."""
    rng = np.random.default_rng(seed)
    dsub = d // M
    cent0 = rng.random((nlist, d), dtype=np.float32)
    x = cent0[rng.integers(0, nlist, n)] + 0.08 * rng.standard_normal((n, d)).astype(np.float32)
    x = (x * scale + shift).astype(np.float32)
    coarse = (cent0 * scale + shift).astype(np.float32)
    list_no = oracle.C.assign(x, coarse)
    res = x - coarse[list_no]
    pq = np.stack([res[rng.integers(0, n, 256), m * dsub:(m + 1) * dsub] for m in range(M)]).astype(np.float32)
    codes = oracle.C.encode(x, coarse, list_no, pq)
    order = np.argsort(list_no, kind="stable")
    offsets = np.zeros(nlist + 1, np.int64)
    offsets[1:] = np.cumsum(np.bincount(list_no, minlength=nlist))
    a = {"coarse": coarse, "pq": np.ascontiguousarray(pq), "offsets": offsets, "codes": np.ascontiguousarray(codes[order]),
         "ids": np.arange(n, dtype=np.int64)[order].copy(), "d": d, "nlist": nlist, "M": M}
    xq = (cent0[rng.integers(0, nlist, nq)] + 0.08 * rng.standard_normal((nq, d)).astype(np.float32))
    return a, (xq * scale + shift).astype(np.float32)


@pytest.mark.parametrize("scale,shift", [(1.0, 0.0), (255.0, 0.0), (1.0, 1000.0), (1e-3, 0.0), (1.0, -37.5), (3e4, 1e6)])
def test_qlut_scan_at_any_scale(oracle, scale, shift):
    """Magnitudes and offsets only change how much the filter removes, never the result."""
    a, xq = _scaled(oracle, 77, 128, 20, 16, 40000, 96, scale, shift)
    _check(oracle, a, xq, 6, 10, f"scale {scale} shift {shift}")


def test_qlut_scan_reference_toy_generator(oracle):
    """IVFPQ_random_dataset.py:6-13: uniform data with x[:, 0] += arange(n) / 1000."""
    rng = np.random.default_rng(1234)
    n, d, nlist, M, nq = 50000, 64, 16, 16, 50
    x = rng.random((n, d)).astype(np.float32)
    x[:, 0] += np.arange(n) / 1000.0
    xq = rng.random((nq, d)).astype(np.float32)
    xq[:, 0] += np.arange(nq) / 1000.0
    coarse = x[rng.choice(n, nlist, replace=False)].copy()
    list_no = oracle.C.assign(x, coarse)
    res = x - coarse[list_no]
    dsub = d // M
    pq = np.stack([res[rng.integers(0, n, 256), m * dsub:(m + 1) * dsub] for m in range(M)]).astype(np.float32)
    codes = oracle.C.encode(x, coarse, list_no, pq)
    order = np.argsort(list_no, kind="stable")
    offsets = np.zeros(nlist + 1, np.int64)
    offsets[1:] = np.cumsum(np.bincount(list_no, minlength=nlist))
    a = {"coarse": coarse, "pq": np.ascontiguousarray(pq), "offsets": offsets, "codes": np.ascontiguousarray(codes[order]),
         "ids": np.arange(n, dtype=np.int64)[order].copy(), "d": d, "nlist": nlist, "M": M}
    _check(oracle, a, xq, 8, 10, "toy generator")


def test_qlut_ties_follow_scan_order(oracle):
    rng = np.random.default_rng(3)
    a = _util.make_index_arrays(oracle, 6, 64, 8, 16, 4000)
    a["codes"] = np.ascontiguousarray(a["codes"][rng.integers(0, 40, size=a["codes"].shape[0])])   # 40 distinct codes
    xq = _util.make_queries(2, a, 30)
    Dr, _ = oracle.C.search(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], 8, 50)
    assert any(len(set(_util.bits(r).tolist())) < 50 for r in Dr), "test needs real ties"
    _check(oracle, a, xq, 8, 50, "ties")


def test_qlut_degenerate_inputs(oracle):
    """A zero PQ codebook (every table entry equal: scale 0), queries equal to a centroid, and a query of zeros."""
    a = _util.make_index_arrays(oracle, 9, 64, 6, 16, 3000)
    xq = _util.make_queries(4, a, 12)
    xq[0] = a["coarse"][2]
    xq[1] = 0.0
    _check(oracle, a, xq, 6, 10, "centroid / zero query")
    a["pq"] = np.zeros_like(a["pq"])
    _check(oracle, a, xq, 6, 10, "zero codebook")


def test_qlut_filter_actually_filters(oracle):
    """The exact path alone would also pass the parity tests: make sure the filter does the work."""
    import b200ivfpq as faiss
    a, xq = _scaled(oracle, 5, 128, 16, 16, 120000, 256, 1.0, 0.0)
    index = _load(a, stats=True, stream=0)
    index.nprobe = 8
    D, I = index.search(xq, 10)
    Dr, Ir = oracle.C.search(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], 8, 10)
    _util.assert_bit_equal(D, Dr, "D")
    _util.assert_bit_equal(I, Ir, "I")
    out = (ctypes.c_int64 * 3)()
    h = index._ensure_handle()
    faiss._lib.check(h.lib.b200_ivfpq_get_filter_stats(h.h, out, 1))
    codes = index.last_scan_stats()["codes"]
    assert out[2] > 0 and out[1] > 0
    assert out[1] < 0.1 * codes, f"{out[1]} exact evaluations for {codes} (query, code) pairs: the filter is not filtering"


@pytest.mark.parametrize("kind", ["1", "2"])
@pytest.mark.parametrize("d,nlist,n,nq,nprobe,k,used", [
    (128, 24, 60000, 103, 5, 10, None),     # several tiles per list (ring of bulk-async tiles wraps), odd group sizes
    (96, 16, 12000, 64, 16, 100, 13),       # dsub 6, k = 100, empty lists
    (64, 8, 3000, 1, 8, 10, None),          # one query: every work item is a single
    (128, 4, 9000, 200, 4, 1, None),        # k = 1, heavy ties
    (128, 6, 300, 20, 6, 10, None),         # lists shorter than one tile
])
def test_two_query_filter_kernel(oracle, d, nlist, n, nq, nprobe, k, used, kind):
    """Work items of two pairs and 32-bit table words, forced for every batch.  kind 1: st_filter_kernel<16, true> (the
    register-pipelined kernel `auto` picks when a list is probed by at most one query on average); kind 2:
    st_filter2_kernel (cp.async.bulk code tiles on an mbarrier ring; opt-in experiment)."""
    a = _util.make_index_arrays(oracle, 290 + d, d, nlist, 16, n, used_lists=used)
    if nlist == 4:
        rng = np.random.default_rng(5)
        a["codes"] = np.ascontiguousarray(a["codes"][rng.integers(0, 500, size=a["codes"].shape[0])])
    xq = _util.make_queries(19, a, nq)
    Dr, Ir = oracle.C.search(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], nprobe, k)
    index = _load(a, env={"B200_IVFPQ_STREAM_TWO": kind, "B200_IVFPQ_SCAN": "qlut"})
    index.nprobe = nprobe
    for _ in range(3):
        D, I = index.search(xq, k)
        _util.assert_bit_equal(D, Dr, "D (two-query filter)")
        _util.assert_bit_equal(I, Ir, "I (two-query filter)")


def test_prepared_query_tables(oracle):
    """b200_ivfpq_prepare_queries: tables built ahead of the search (same queries) give the oracle's result; a prepare for
    OTHER queries is ignored; a prepare before the first search of a handle (buffers not sized yet) is a no-op."""
    import torch
    a, xq = _scaled(oracle, 33, 128, 16, 16, 50000, 96, 1.0, 0.0)
    nprobe, k, nq = 6, 10, xq.shape[0]
    Dr, Ir = oracle.C.search(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], nprobe, k)
    index = _load(a)
    index.nprobe = nprobe
    xq_t = torch.from_numpy(xq).cuda()
    other = torch.flip(xq_t, dims=(0,)).contiguous()
    for step in range(4):
        held = index.prepare_queries(xq_t if step != 2 else other)      # step 0: before the first search
        _, probes = index.quantizer.search(xq_t, nprobe)
        if step == 3:
            D, I = index.search_preassigned(held, k, probes)
        else:
            thr = index.search_preassigned_begin(xq_t if step == 2 else held, k, probes, 0, nq)
            assert thr is not None
            D, I = index.search_preassigned_finish(thr, nq, k)
        _util.assert_bit_equal(D.cpu().numpy(), Dr, f"D (prepared tables, step {step})")
        _util.assert_bit_equal(I.cpu().numpy(), Ir, f"I (prepared tables, step {step})")
    # and the plain search right after a prepare of the same tensor
    held = index.prepare_queries(xq_t)
    D, I = index.search(held, k)
    _util.assert_bit_equal(D.cpu().numpy(), Dr, "D (prepared, full search)")
    _util.assert_bit_equal(I.cpu().numpy(), Ir, "I (prepared, full search)")


def test_split_search_begin_finish(oracle):
    """b200_ivfpq_search_preassigned_begin / _finish (the multi-GPU threshold exchange) on one GPU: thresholds of all
    queries, of half of them (the other half then has none: everything survives, the buffers overflow, the fallback
    answers), and thresholds tightened by a second index holding the same vectors -- always the oracle's result."""
    import torch
    a, xq = _scaled(oracle, 31, 128, 16, 16, 50000, 96, 1.0, 0.0)
    nprobe, k, nq = 6, 10, xq.shape[0]
    Dr, Ir = oracle.C.search(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], nprobe, k)
    index = _load(a)
    index.nprobe = nprobe
    xq_t = torch.from_numpy(xq).cuda()
    _, probes = index.quantizer.search(xq_t, nprobe)
    for lo, hi in ((0, nq), (0, nq // 2), (nq // 3, nq)):
        thr = index.search_preassigned_begin(xq_t, k, probes, lo, hi)
        assert thr is not None, "the streaming pipeline applies to this shape"
        t = thr.cpu().numpy()
        assert (t[:lo] == 0x7f800000).all() and (t[hi:] == 0x7f800000).all() and (t[lo:hi] < 0x7f800000).all()
        D, I = index.search_preassigned_finish(thr, nq, k)
        _util.assert_bit_equal(D.cpu().numpy(), Dr, f"D (slice {lo}:{hi})")
        _util.assert_bit_equal(I.cpu().numpy(), Ir, f"I (slice {lo}:{hi})")
    # thresholds from elsewhere (here: the true k-th distances, the tightest valid ones) are applied by distance only
    thr = index.search_preassigned_begin(xq_t, k, probes, 0, 0)
    best = torch.from_numpy(np.ascontiguousarray(Dr[:, k - 1])).cuda().view(torch.int32)
    D, I = index.search_preassigned_finish(torch.minimum(thr, best), nq, k)
    _util.assert_bit_equal(D.cpu().numpy(), Dr, "D (exchanged thresholds)")
    _util.assert_bit_equal(I.cpu().numpy(), Ir, "I (exchanged thresholds)")


def test_stream_overflow_falls_back(oracle):
    """A survivor buffer far too small for the batch: the device raises the overflow flag and the guarded launches
    (in-kernel path + merge) answer; the results are the oracle's either way."""
    import b200ivfpq as faiss
    a, xq = _scaled(oracle, 11, 128, 16, 16, 60000, 128, 1.0, 0.0)
    Dr, Ir = oracle.C.search(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], 8, 10)
    for minrec, expect_overflow in (("64", True), ("4194304", False)):
        index = _load(a, stats=True, env={"B200_IVFPQ_STREAM_MINREC": minrec, "B200_IVFPQ_STREAM_RATE": "1e-6"})
        index.nprobe = 8
        D, I = index.search(xq, 10)
        _util.assert_bit_equal(D, Dr, f"D (minrec {minrec})")
        _util.assert_bit_equal(I, Ir, f"I (minrec {minrec})")
        st = index.filter_stats(reset=True)
        assert (st["work_items"] < 0) == expect_overflow, st


def test_stream_survivors_are_few(oracle):
    """With bootstrap thresholds the filter passes well under 2 % of the (query, code) pairs on clustered data."""
    a, xq = _scaled(oracle, 5, 128, 16, 16, 120000, 256, 1.0, 0.0)
    index = _load(a, stats=True)
    index.nprobe = 8
    D, I = index.search(xq, 10)
    Dr, Ir = oracle.C.search(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], 8, 10)
    _util.assert_bit_equal(D, Dr, "D")
    _util.assert_bit_equal(I, Ir, "I")
    st = index.filter_stats(reset=True)
    codes = index.last_scan_stats()["codes"]
    assert st["work_items"] >= 0, "unexpected overflow"
    assert 0 < st["exact_evaluations"] < 0.02 * codes, (st, codes)


def test_qlut_lists_replaced(oracle):
    """set_lists again (index.add after a search): the per-vector term is rebuilt."""
    a = _util.make_index_arrays(oracle, 21, 64, 8, 16, 6000)
    xq = _util.make_queries(5, a, 40)
    index = _check(oracle, a, xq, 4, 10, "first lists")
    b = _util.make_index_arrays(oracle, 22, 64, 8, 16, 9000)
    index.set_codebooks(b["coarse"], b["pq"])
    index.set_lists(b["offsets"], b["codes"], b["ids"])
    Dr, Ir = oracle.C.search(xq, b["coarse"], b["pq"], b["offsets"], b["codes"], b["ids"], 4, 10)
    D, I = index.search(xq, 10)
    _util.assert_bit_equal(D, Dr, "D (second lists)")
    _util.assert_bit_equal(I, Ir, "I (second lists)")
