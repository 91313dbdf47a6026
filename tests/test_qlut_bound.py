"""The lower bound of the per-query-table filter (csrc/scan_qlut.cuh, csrc/scan_stream.cuh), restated in numpy and
checked on the CPU: a code whose ORACLE distance is at or below the threshold must never be dropped,

    sum_m u_q[m][code_m]  <=  floor-free float test of  s_q (thr + E - dis0 + sum_m B_m - sbmin - sbstep v)  + slack

for every query / list / code, at unit scale, at SIFT scale, far from the origin (where the decomposition
||q-c-p||^2 = ||q-c||^2 + [||p||^2 + 2(c-mu).p] + [-2(q-mu).p] cancels badly) and on the reference's toy generator
with its 0..n/1000 ramp.  The kernels evaluate the same formulas in the same fp32 operations (constants included);
GPU parity of the complete path is in tests/test_gpu_qlut.py.
"""
import numpy as np
import pytest

f32 = np.float32
U = f32(5.9604645e-8)       # 2^-24


def oracle_distances(q, c, pq, codes):
    """sum_m sum_j ((q - c)[m*dsub+j] - pq[m][code_m][j])^2, j ascending then m ascending, every op rounded to fp32."""
    M, _, dsub = pq.shape
    r = (q - c).astype(f32)
    acc = np.zeros(codes.shape[0], f32)
    for m in range(M):
        t = np.zeros(codes.shape[0], f32)
        p = pq[m][codes[:, m]]                                   # (n, dsub)
        for j in range(dsub):
            diff = (r[m * dsub + j] - p[:, j]).astype(f32)
            t = (t + (diff * diff).astype(f32)).astype(f32)
        acc = (acc + t).astype(f32)
    return acc


def query_table(q, mu, pq, pqmax, qmax):
    """ql_query_tables_kernel: offsets / scale from the Cauchy-Schwarz bound, entries floor((A + B_m) s (1 - 2^-20))."""
    M, _, dsub = pq.shape
    qc = (q - mu).astype(f32)
    B = np.empty(M, f32)
    for m in range(M):
        nn = f32(np.sum(qc[m * dsub:(m + 1) * dsub].astype(np.float64) ** 2))
        B[m] = f32(2.0) * np.sqrt(nn).astype(f32) * f32(1.0001) * pqmax[m]
    rng = f32(2.0) * B.max()
    sumB = f32(B.sum())
    s = f32(qmax) / (rng * f32(1.0001)) if rng > 0 else f32(0)
    slack = f32(dsub + 3) * f32(1.1920929e-7) * sumB + sumB * f32(1.2e-7) * f32(M)
    amin = -(sumB + slack)
    A = (f32(-2.0) * np.einsum("mkj,mj->mk", pq, qc.reshape(M, dsub)).astype(f32)).astype(f32)
    x = ((A + B[:, None]) * (s * f32(0.999999))).astype(f32)
    u = np.where(x > 0, np.minimum(np.floor(x), qmax), 0).astype(np.int64)
    return u, s, f32(amin)


def vector_term(c, mu, pq, codes):
    """ql_sb_build_kernel: SB in float64, 16 bits on the list's grid, rounded down."""
    M, _, dsub = pq.shape
    G = (pq.astype(np.float64) * (pq.astype(np.float64) + 2.0 * (c.astype(np.float64) - mu.astype(np.float64)).reshape(M, 1, dsub))).sum(2)
    sb = G[np.arange(M)[None, :], codes].sum(1)
    lo, hi = sb.min(), sb.max()
    fmin = f32(lo)
    if float(fmin) > lo:
        fmin = np.nextafter(fmin, f32(-np.inf))
    span = hi - float(fmin)
    st = f32(span / 65535.0)
    while float(st) * 65535.0 < span:
        st = np.nextafter(st, f32(np.inf))
    v = np.floor((sb - float(fmin)) / float(st)).astype(np.int64) if st > 0 else np.zeros(len(sb), np.int64)
    v = np.clip(v, 0, 65535)
    v -= (float(fmin) + float(st) * v > sb)
    assert (float(fmin) + float(st) * v <= sb).all()
    return v, fmin, st


def passes(u_sum, v, thr, s, amin, dis0_sum, fmin, st, pmax, d, M, dsub):
    """st_filter_kernel / scan_qlut_kernel: the pair constants and the magic-number comparison."""
    dis0 = dis0_sum * (f32(1.0) - f32(d + 5) * U)
    rn = np.sqrt(dis0_sum).astype(f32) * f32(1.00001) + pmax
    E = f32(dsub + M + 8) * U * rn * rn * f32(1.00001)
    mag = abs(E) + abs(dis0) + abs(amin) + abs(fmin)
    base = f32(f32(f32(E - dis0) - amin) - fmin) + f32(4.8e-7) * mag
    na = -(s * st * f32(0.999999))
    y = f32(thr + base) + f32(4.8e-7) * (abs(thr) + mag)
    sy = f32(s * y)
    tb = f32(sy + abs(sy) * f32(4.8e-7)) + f32(2.0 + 8388608.0)
    z = (v.astype(f32) * na + tb).astype(f32)          # the kernel's single-rounding FMA is at least as accurate
    f = (f32(8388608.0) + u_sum.astype(f32)).astype(f32)
    return ~(f > z)


def make_case(seed, d, M, nlist, n, scale, shift, toy=False):
    rng = np.random.default_rng(seed)
    dsub = d // M
    if toy:
        x = rng.random((n, d)).astype(f32)
        x[:, 0] += np.arange(n, dtype=f32) / f32(1000.0)
        cent = x[rng.choice(n, nlist, replace=False)].copy()
    else:
        cent0 = rng.random((nlist, d)).astype(f32)
        x = ((cent0[rng.integers(0, nlist, n)] + 0.08 * rng.standard_normal((n, d))) * scale + shift).astype(f32)
        cent = (cent0 * scale + shift).astype(f32)
    ln = ((x[:, None, :].astype(np.float64) - cent[None].astype(np.float64)) ** 2).sum(2).argmin(1)
    res = x - cent[ln]
    pq = np.stack([res[rng.integers(0, n, 256), m * dsub:(m + 1) * dsub] for m in range(M)]).astype(f32)
    codes = np.stack([((res[:, None, m * dsub:(m + 1) * dsub] - pq[m][None]) ** 2).sum(2).argmin(1) for m in range(M)], 1)
    return cent, pq, ln, codes.astype(np.int64), x


@pytest.mark.parametrize("scale,shift,toy", [(1.0, 0.0, False), (255.0, 0.0, False), (1.0, 1000.0, False),
                                             (1e-3, 0.0, False), (3e4, 1e6, False), (1.0, 0.0, True)])
@pytest.mark.parametrize("M,qmax", [(16, 2047), (32, 2047)])
def test_filter_never_drops_a_result(scale, shift, toy, M, qmax):
    d, nlist, n = 64, 6, 3000
    dsub = d // M
    cent, pq, ln, codes, x = make_case(11, d, M, nlist, n, scale, shift, toy)
    mu = cent.astype(np.float64).mean(0).astype(f32)
    pqmax = (np.sqrt((pq.astype(np.float64) ** 2).sum(2)).max(1) * 1.00001).astype(f32)
    pmax = f32(np.sqrt((pqmax.astype(np.float64) ** 2).sum()) * 1.0001)
    rng = np.random.default_rng(5)
    checked = kept = 0
    for qi in rng.integers(0, n, 12):
        q = (x[qi] + f32(0.01 * scale) * rng.standard_normal(d).astype(f32)).astype(f32)
        u, s, amin = query_table(q, mu, pq, pqmax, qmax)
        for l in range(nlist):
            sel = np.nonzero(ln == l)[0]
            if len(sel) == 0:
                continue
            cc = codes[sel]
            exact = oracle_distances(q, cent[l], pq, cc)
            v, fmin, st = vector_term(cent[l], mu, pq, cc)
            u_sum = u[np.arange(M)[None, :], cc].sum(1)
            r = (q - cent[l]).astype(f32)
            dis0_sum = f32(np.sum(r.astype(np.float64) ** 2))
            for thr in (np.sort(exact)[min(9, len(exact) - 1)], np.median(exact), exact.min()):
                ok = passes(u_sum, v, f32(thr), s, amin, dis0_sum, fmin, st, pmax, d, M, dsub)
                must = exact <= thr
                assert ok[must].all(), f"dropped {int((must & ~ok).sum())} codes at or below the threshold"
                checked += int(must.sum())
                kept += int(ok.sum())
    assert checked > 0


def test_filter_is_selective_at_unit_scale():
    """The bound is not vacuous: with the 10th-best distance as the threshold, most codes are dropped."""
    d, M, nlist, n, qmax = 64, 16, 4, 8000, 2047
    dsub = d // M
    cent, pq, ln, codes, x = make_case(3, d, M, nlist, n, 1.0, 0.0)
    mu = cent.astype(np.float64).mean(0).astype(f32)
    pqmax = (np.sqrt((pq.astype(np.float64) ** 2).sum(2)).max(1) * 1.00001).astype(f32)
    pmax = f32(np.sqrt((pqmax.astype(np.float64) ** 2).sum()) * 1.0001)
    q = x[17]
    u, s, amin = query_table(q, mu, pq, pqmax, qmax)
    l = int(ln[17])
    cc = codes[ln == l]
    exact = oracle_distances(q, cent[l], pq, cc)
    v, fmin, st = vector_term(cent[l], mu, pq, cc)
    u_sum = u[np.arange(M)[None, :], cc].sum(1)
    r = (q - cent[l]).astype(f32)
    ok = passes(u_sum, v, f32(np.sort(exact)[9]), s, amin, f32(np.sum(r.astype(np.float64) ** 2)), fmin, st, pmax, d, M, dsub)
    assert ok[exact <= np.sort(exact)[9]].all()
    assert ok.mean() < 0.2, f"{ok.mean():.2%} of the list survives a top-10 threshold"
