"""C oracle vs its numpy twin (written from the notebook functions), edge cases included."""
import numpy as np
import pytest

import _util


@pytest.mark.parametrize("d,nlist,M,n,nq,nprobe,k", [
    (32, 16, 8, 2000, 12, 4, 10),
    (24, 10, 4, 500, 8, 10, 100),     # nprobe == nlist, k large
    (16, 32, 16, 300, 6, 5, 1),       # dsub = 1
    (96, 8, 16, 900, 5, 3, 7),        # dsub = 6 (Deep1B shape)
    (20, 6, 5, 400, 4, 2, 3),         # M not a power of two
])
def test_search_matches_numpy_twin(oracle, d, nlist, M, n, nq, nprobe, k):
    a = _util.make_index_arrays(oracle, 11, d, nlist, M, n)
    xq = _util.make_queries(3, a, nq)
    D, I, pdis, pid = oracle.C.search(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], nprobe, k,
                                      return_probes=True)
    ndis, nid = oracle.np_coarse(xq, a["coarse"], nprobe)
    _util.assert_bit_equal(pid, nid, "probe ids")
    _util.assert_bit_equal(pdis, ndis, "probe distances")
    D2, I2 = oracle.np_search(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], nprobe, k)
    _util.assert_bit_equal(D, D2, "D")
    _util.assert_bit_equal(I, I2, "I")
    assert (np.diff(D.astype(np.float64), axis=1) >= 0).all()


def test_lut_adc_match_notebook_formulas(oracle):
    rng = np.random.default_rng(0)
    d, M = 32, 8
    q, c = rng.random(d, dtype=np.float32), rng.random(d, dtype=np.float32)
    pq = rng.standard_normal((M, 256, d // M)).astype(np.float32)
    T = oracle.C.lut(q, c, pq)
    _util.assert_bit_equal(T, oracle.np_lut(q, c, pq), "LUT")
    # construct_distance_table (ipynb:7929-7946) in float64 agrees to fp32 rounding
    res = (q - c).astype(np.float64).reshape(M, 1, -1)
    np.testing.assert_allclose(T, ((res - pq.astype(np.float64)) ** 2).sum(2), rtol=1e-5)
    codes = rng.integers(0, 256, size=(500, M), dtype=np.uint8)
    dist = oracle.C.adc(T, codes)
    _util.assert_bit_equal(dist, oracle.np_adc(T, codes), "ADC")
    # estimate_distance (ipynb:7948-7960), python loop, float64 accumulate
    for i in (0, 17, 499):
        ref = sum(float(T[m, codes[i, m]]) for m in range(M))
        assert abs(dist[i] - ref) <= 1e-5 * ref


def test_l2sqr_is_sequential_non_fused(oracle):
    """The contract: each product and each add rounded to fp32 separately, j ascending."""
    rng = np.random.default_rng(1)
    a, b = rng.standard_normal(77).astype(np.float32), rng.standard_normal(77).astype(np.float32)
    acc = np.float32(0)
    for j in range(77):
        diff = np.float32(a[j] - b[j])
        acc = np.float32(acc + np.float32(diff * diff))
    assert oracle.C.l2sqr(a, b).view(np.uint32) == acc.view(np.uint32)


def test_empty_and_ragged_lists(oracle):
    a = _util.make_index_arrays(oracle, 2, 16, 12, 4, 150, used_lists=5)
    sizes = np.diff(a["offsets"])
    assert (sizes == 0).sum() >= 1
    xq = _util.make_queries(9, a, 10)
    D, I = oracle.C.search(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], 12, 20)
    D2, I2 = oracle.np_search(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], 12, 20)
    _util.assert_bit_equal(D, D2)
    _util.assert_bit_equal(I, I2)


def test_unfilled_slots(oracle):
    """Fewer candidates than k: id -1 and distance FLT_MAX (Faiss 1.7.1 convention)."""
    a = _util.make_index_arrays(oracle, 4, 8, 6, 2, 9)
    xq = _util.make_queries(1, a, 3)
    D, I = oracle.C.search(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], 1, 16)
    assert (I == -1).any()
    assert (D[I == -1] == oracle.FLT_MAX).all()
    assert (D[I != -1] < oracle.FLT_MAX).all()
    # empty index
    off0 = np.zeros(7, np.int64)
    D, I = oracle.C.search(xq, a["coarse"], a["pq"], off0, np.zeros((0, 2), np.uint8), np.zeros(0, np.int64), 3, 4)
    assert (I == -1).all() and (D == oracle.FLT_MAX).all()


def test_tie_rule_earlier_scanned_wins(oracle):
    """Duplicate codes: equal distances.  Strict '<' replace (priority_queue_L1.hpp:65-75) = the earlier-scanned
    entry (probe rank, then list offset) ranks first and survives at the k boundary."""
    rng = np.random.default_rng(8)
    d, M, nlist = 8, 4, 3
    coarse = rng.random((nlist, d), dtype=np.float32)
    pq = rng.standard_normal((M, 256, 2)).astype(np.float32)
    code = rng.integers(0, 256, size=(1, M), dtype=np.uint8)
    codes = np.repeat(code, 12, axis=0)                     # 12 identical codes, 4 per list
    offsets = np.array([0, 4, 8, 12], np.int64)
    ids = np.arange(100, 112, dtype=np.int64)
    xq = coarse[:1].copy()                                   # nearest list is 0
    pid = np.array([[0]], np.int64)
    D, I = oracle.C.search_preassigned(xq, coarse, pq, offsets, codes, ids, pid, 3)
    assert I.tolist() == [[100, 101, 102]]
    assert len(set(_util.bits(D).ravel().tolist())) == 1
    D2, I2 = oracle.np_search_preassigned(xq, coarse, pq, offsets, codes, ids, pid, 3)
    _util.assert_bit_equal(I, I2)


def test_preassigned_skips_negative_lists(oracle):
    a = _util.make_index_arrays(oracle, 6, 16, 8, 4, 400)
    xq = _util.make_queries(2, a, 5)
    pid = np.array([[0, -1, 3], [-1, -1, -1], [7, 7, 1], [2, 5, -1], [4, 0, 6]], np.int64)
    D, I = oracle.C.search_preassigned(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], pid, 5)
    D2, I2 = oracle.np_search_preassigned(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], pid, 5)
    _util.assert_bit_equal(D, D2)
    _util.assert_bit_equal(I, I2)
    assert (I[1] == -1).all()


def test_coarse_ties_lower_id(oracle):
    cent = np.zeros((6, 4), np.float32)
    cent[3] = 1.0
    xq = np.zeros((1, 4), np.float32)
    dis, ids = oracle.C.coarse(xq, cent, 4)
    assert ids.tolist() == [[0, 1, 2, 4]]
    _, ids2 = oracle.np_coarse(xq, cent, 4)
    assert ids2.tolist() == ids.tolist()


def test_encode_is_argmin_of_lut(oracle):
    a = _util.make_index_arrays(oracle, 12, 16, 5, 4, 60)
    x, ln = a["x"], a["list_no"]
    codes = oracle.C.encode(x, a["coarse"], ln, a["pq"])
    for i in (0, 31, 59):
        T = oracle.C.lut(x[i], a["coarse"][ln[i]], a["pq"])
        assert codes[i].tolist() == T.argmin(axis=1).tolist()
    # assign = nprobe-1 coarse search
    _, ids = oracle.C.coarse(x, a["coarse"], 1)
    assert np.array_equal(ids[:, 0], ln)


def test_merge_shards_equals_single_index(oracle):
    """Shard by position modulo (gpu_recall:3-7), search each shard, merge = concatenate + argsort
    (bench_multi_cpu_performance_OSDI.py:203-219) -> same distances as the unsharded search."""
    a = _util.make_index_arrays(oracle, 21, 32, 16, 8, 3000, id_scramble=False)
    xq = _util.make_queries(5, a, 20)
    nprobe, k, world = 6, 10, 3
    D, I = oracle.C.search(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], nprobe, k)
    Ds, Is = [], []
    sizes = np.diff(a["offsets"])
    list_no = np.repeat(np.arange(16), sizes)
    for r in range(world):
        keep = (a["ids"] % world) == r
        off = np.zeros(17, np.int64)
        off[1:] = np.cumsum(np.bincount(list_no[keep], minlength=16))
        d_, i_ = oracle.C.search(xq, a["coarse"], a["pq"], off, a["codes"][keep], a["ids"][keep], nprobe, k)
        Ds.append(d_)
        Is.append(i_)
    Dm, Im = oracle.C.merge_shards(np.stack(Ds), np.stack(Is))
    _util.assert_same_modulo_ties(Dm, Im, D, I, "merged vs single")
    Dn, In = oracle.np_merge_shards(np.stack(Ds), np.stack(Is))
    _util.assert_bit_equal(Dm, Dn)
    _util.assert_bit_equal(Im, In)


def test_recall_formulas(oracle):
    """bench_gpu_performance_OSDI.py:196-200, 690-692."""
    I = np.array([[1, 2, 3], [4, 5, 6]])
    gt = np.array([[2, 9, 1], [7, 4, 8]])
    assert oracle.recall_at_k(I, gt, 3) == pytest.approx(3 / 6)
    assert oracle.r1_at_k(I, gt, 3) == pytest.approx(1 / 2)
    assert oracle.r1_at_k(I, gt, 1) == 0.0


@pytest.mark.parametrize("scale,sigma", [(1.0, 0.05), (1.0, 0.01), (255.0, 0.02)])
def test_faiss_precomputed_table_form_stays_inside_the_stated_tolerance(oracle, scale, sigma):
    """Faiss-CPU evaluates dis0 + sum_m (P[l][m][code] - 2 <q_m, p>) (use_precomputed_table = 1), the
    contract here is the notebook's residual-LUT form.  The two are not bit-identical; north_star's tolerance (distances
    within 1e-5 relative, ids equal except ties) has to cover the gap -- measured here on clustered data at SIFT-like
    (0..255) and unit scales."""
    rng = np.random.default_rng(0)
    d, nlist, M, n, nq, nprobe, k = 128, 32, 16, 6000, 24, 8, 10
    coarse = (rng.random((nlist, d), dtype=np.float32) * scale).astype(np.float32)
    pq = (rng.standard_normal((M, 256, d // M)).astype(np.float32) * sigma * scale).astype(np.float32)
    off = np.zeros(nlist + 1, np.int64)
    off[1:] = np.cumsum(rng.multinomial(n, np.full(nlist, 1 / nlist)))
    codes = rng.integers(0, 256, (n, M), dtype=np.uint8)
    ids = np.arange(n, dtype=np.int64)
    xq = (coarse[rng.integers(0, nlist, nq)] + rng.standard_normal((nq, d)).astype(np.float32) * sigma * scale)
    xq = xq.astype(np.float32)
    _, pid = oracle.C.coarse(xq, coarse, nprobe)
    D, I = oracle.C.search_preassigned(xq, coarse, pq, off, codes, ids, pid, k)
    D2, I2 = oracle.np_search_preassigned_precomputed(xq, coarse, pq, off, codes, ids, pid, k)
    rel = np.abs(D - D2) / np.abs(D)
    assert rel.max() < 1e-5, rel.max()
    assert rel.max() > 0, "the two forms are expected to differ in the last bits"
    same_sets = np.mean([len(set(a) & set(b)) / k for a, b in zip(I, I2)])
    assert same_sets >= 0.99
