"""Shared helpers for the tests: seeded synthetic indexes built WITH THE ORACLE, and result comparison."""
import numpy as np

GOLDEN = __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "golden")


def make_index_arrays(oracle, seed, d, nlist, M, n, used_lists=None, id_scramble=True, sigma=0.15):
    """Random codebooks + n vectors assigned and encoded by the oracle, in the flattened
    ArrayInvertedLists layout.  used_lists < nlist leaves the other lists empty (ragged / empty case)."""
    rng = np.random.default_rng(seed)
    dsub = d // M
    coarse = rng.random((nlist, d), dtype=np.float32)
    pq = (rng.standard_normal((M, 256, dsub)) * sigma).astype(np.float32)
    live = nlist if used_lists is None else used_lists
    which = rng.integers(0, live, size=n)
    x = (coarse[which] + rng.standard_normal((n, d)).astype(np.float32) * sigma).astype(np.float32)
    list_no = oracle.C.assign(x, coarse) if n else np.zeros(0, np.int64)
    codes = oracle.C.encode(x, coarse, list_no, pq) if n else np.zeros((0, M), np.uint8)
    ids = (rng.permutation(n).astype(np.int64) * 7 + 3) if id_scramble else np.arange(n, dtype=np.int64)
    order = np.argsort(list_no, kind="stable")
    offsets = np.zeros(nlist + 1, np.int64)
    offsets[1:] = np.cumsum(np.bincount(list_no, minlength=nlist))
    return {"coarse": coarse, "pq": pq, "offsets": offsets, "codes": np.ascontiguousarray(codes[order]),
            "ids": np.ascontiguousarray(ids[order]), "x": x, "list_no": list_no, "d": d, "nlist": nlist, "M": M}


def make_queries(seed, arrays, nq, sigma=0.15):
    rng = np.random.default_rng(seed)
    c = arrays["coarse"]
    which = rng.integers(0, c.shape[0], size=nq)
    return (c[which] + rng.standard_normal((nq, c.shape[1])).astype(np.float32) * sigma).astype(np.float32)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_bit_equal(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if a.dtype.kind == "f":
        bad = bits(a) != bits(b)
    else:
        bad = a != b
    assert not bad.any(), f"{what}: {int(bad.sum())} of {bad.size} entries differ; first at {np.argwhere(bad)[0]}"


def assert_same_modulo_ties(D, I, D_ref, I_ref, what=""):
    """Distances bit-exact; ids identical except inside runs of equal distance, where the id SETS must match
    (for the run that touches the k-th slot, ids only need to be drawn from candidates of that distance, which
    the caller guarantees by construction or checks separately)."""
    assert_bit_equal(D, D_ref, what + " distances")
    nq, k = D.shape
    for q in range(nq):
        if np.array_equal(I[q], I_ref[q]):
            continue
        i = 0
        while i < k:
            j = i
            while j + 1 < k and bits(D_ref[q, j + 1]) == bits(D_ref[q, i]):
                j += 1
            a, b = sorted(I[q, i:j + 1].tolist()), sorted(I_ref[q, i:j + 1].tolist())
            if a != b:
                assert j == k - 1, f"{what}: query {q} ids differ outside a boundary tie: {I[q]} vs {I_ref[q]}"
            i = j + 1


def assert_same_modulo_near_ties(D, I, D_ref, I_ref, rtol, what=""):
    """For comparisons against code that sums in another order (the reference's SIMD kernels): the sorted distance
    profiles agree within `rtol`, and wherever the ids differ the two entries involved are a near tie -- their distances
    differ by at most 2*rtol relative (an order swap), or the missing entry sat within 2*rtol of the k-th distance."""
    D, D_ref = np.asarray(D, np.float32), np.asarray(D_ref, np.float32)
    np.testing.assert_allclose(D, D_ref, rtol=rtol, atol=0, err_msg=what + " distances")
    nq, k = D.shape
    for q in range(nq):
        if np.array_equal(I[q], I_ref[q]):
            continue
        mine = {int(i): p for p, i in enumerate(I[q])}
        for pos in np.nonzero(I[q] != I_ref[q])[0]:
            other = int(I_ref[q, pos])
            d_here = float(D_ref[q, pos])
            d_there = float(D[q, mine[other]]) if other in mine else float(D[q, k - 1])
            assert abs(d_there - d_here) <= 2 * rtol * max(d_here, 1e-30), \
                f"{what}: query {q} position {pos}: id {other} vs {int(I[q, pos])} is not a near tie"
