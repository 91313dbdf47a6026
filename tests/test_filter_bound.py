"""The four-query scan kernel (csrc/scan_quad.cuh) drops a code without evaluating it exactly when an integer lower
bound of its distance exceeds an integer image of the current threshold.  This file restates that arithmetic in numpy
(fp32 where the kernel uses fp32) and checks the property the kernel's exactness rests on, on the CPU:

    exact_fp32(code) <= thr   ==>   LB_int(code) <= t_int(thr)            (no true result is ever dropped)

together with the no-overflow condition of the packed 16-bit accumulators.  exact_fp32 is the oracle's ADC: 16
sequential fp32 adds in ascending m (IVFPQ_1B_search.ipynb:7948-7960).
"""
import numpy as np
import pytest

F = np.float32


def quantised_lut(T, r_norms, p_maxnorm):
    """scan_quad.cuh: B = max_m (||r_m|| * 1.00001 + pmax_m)^2 * 1.0001, s = 2047 / B * 0.999999,
    u[m][c] = min(trunc(T[m][c] * s), 2047), everything in fp32."""
    b = (r_norms.astype(F) * F(1.00001) + p_maxnorm.astype(F)) ** 2 * F(1.0001)
    B = b.max().astype(F)
    s = (F(2047.0) / B) * F(0.999999) if B > 0 else F(0.0)
    u = np.minimum(np.trunc(T.astype(F) * s).astype(np.int64), 2047)
    return u, s


def int_threshold(thr, s):
    """quad_int_threshold: inf -> 0x7fff; else trunc(thr * s * 1.000004) + 1, saturating at 0x7fff."""
    if not np.isfinite(thr):
        return 0x7FFF
    x = F(thr) * s * F(1.000004)
    if not x < F(32000.0):
        return 0x7FFF
    return int(np.trunc(x)) + 1


def exact_fp32(T, codes):
    acc = np.zeros(codes.shape[0], F)
    for m in range(T.shape[0]):
        acc = (acc + T[m, codes[:, m]]).astype(F)          # one rounding per add, ascending m
    return acc


@pytest.mark.parametrize("seed,dsub,spread", [(0, 8, 1.0), (1, 6, 0.05), (2, 8, 30.0), (3, 4, 1e-3), (4, 16, 1.0)])
def test_lower_bound_never_drops_a_result(seed, dsub, spread):
    rng = np.random.default_rng(seed)
    M, n = 16, 200_000
    pq = (rng.standard_normal((M, 256, dsub)) * spread).astype(F)
    r = (rng.standard_normal((M, dsub)) * spread * rng.uniform(0.1, 3.0)).astype(F)
    # LUT exactly as the kernels build it: sum_j (r_j - p_j)^2, sequential, separately rounded
    T = np.zeros((M, 256), F)
    for j in range(dsub):
        diff = (r[:, None, j] - pq[:, :, j]).astype(F)
        T = (T + (diff * diff).astype(F)).astype(F)
    r_norms = np.sqrt((r.astype(np.float64) ** 2).sum(1))
    p_maxnorm = np.sqrt((pq.astype(np.float64) ** 2).sum(2)).max(1) * 1.00001
    u, s = quantised_lut(T, r_norms, p_maxnorm)
    assert (np.trunc(T * s) <= 2047).all(), "B must bound every table entry (the clamp is only a safety net)"
    codes = rng.integers(0, 256, size=(n, M))
    codes[: n // 10] = codes[rng.integers(0, 50, size=n // 10)]      # duplicates: ties at the threshold
    exact = exact_fp32(T, codes)
    lb = u[np.arange(M)[None, :], codes].sum(1)
    assert lb.max() <= 16 * 2047 < 2 ** 15, "packed u16 accumulators must not carry into their neighbour"
    # thresholds: the k-th best distance for several k, plus values exactly equal to some distances
    order = np.sort(exact)
    for thr in [order[0], order[9], order[99], order[999], order[n // 2], order[-1], np.inf]:
        t = int_threshold(thr, s)
        passed = lb <= t
        assert passed[exact <= thr].all(), f"a result with distance <= {thr} was filtered out"
    # and the filter is worth having: at the 10th-best threshold it lets through a tiny fraction
    t = int_threshold(order[9], s)
    assert (lb <= t).mean() < 0.02


def test_guard_bit_compare():
    """(0x8000 | t) - s keeps bit 15 iff s <= t for s, t < 2^15; two queries per 32-bit word without borrow."""
    rng = np.random.default_rng(0)
    t0, t1, s0, s1 = (rng.integers(0, 2 ** 15, size=100_000, dtype=np.int64) for _ in range(4))
    tw = (0x80008000 | (t1 << 16) | t0) & 0xFFFFFFFF
    sw = ((s1 << 16) | s0) & 0xFFFFFFFF
    d = (tw - sw) & 0xFFFFFFFF
    assert np.array_equal((d >> 15) & 1, (s0 <= t0).astype(np.int64))
    assert np.array_equal((d >> 31) & 1, (s1 <= t1).astype(np.int64))
