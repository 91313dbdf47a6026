#!/usr/bin/env python
"""CPU-only design experiment for the per-QUERY lower-bound filter (csrc/scan_qlut.cuh).

The four-query filter of round 1 builds one quantised table per (query, list) pair from the residual.  This experiment
checks the decomposition that needs no per-pair table at all:

    ||q - c - p||^2 = ||q - c||^2 + [ ||p||^2 + 2 (c - mu).p ] + [ -2 (q - mu).p ]
                    =    dis0     +        SB (per stored vector)  +  sum_m A_q[m][code_m] (per query)

with A quantised per query (u = floor((A - min_m) * s_q), s_q from the largest range over m) and SB stored per vector
as a 16-bit (or 8-bit) value on a per-list grid, rounded down.  A code survives when
    sum_m u[m][code_m] <= floor(s_q * (thr - dis0 - sum_m min_m - SB_down)) + 1.
Reports the survivor rate next to the per-pair filter's (tools/filter_precision_experiment.py) on the same index.

    python tools/qlut_filter_experiment.py [--nb 400000] [--nq 60]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "chameleon-rag-acceleration_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from b200ivfpq.datasets import SEED_BASE, SEED_QUERY, SEED_TRAIN, ClusteredGenerator  # noqa: E402
from b200ivfpq.kmeans import kmeans, kmeans_subspaces  # noqa: E402
from filter_precision_experiment import assign, encode  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nb", type=int, default=400000)
    ap.add_argument("--nq", type=int, default=60)
    ap.add_argument("--nlist", type=int, default=128)
    ap.add_argument("--nprobe", type=int, default=16)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--scale", type=float, default=1.0, help="multiply the data (SIFT-like magnitudes: 255)")
    ap.add_argument("--shift", type=float, default=0.0, help="add a constant to every coordinate (uncentred data)")
    a = ap.parse_args()
    d, M = 128, 16
    dsub = d // M
    gen = ClusteredGenerator(d, ncentres=a.nlist // 2, sigma=0.1, device=torch.device("cpu"), seed=7, latent_dim=12,
                             sigma_iso=0.002)
    tf = lambda x: (x * a.scale + a.shift).astype(np.float32)
    xt = tf(gen.chunk(SEED_TRAIN, 0, 40000).numpy())
    coarse = kmeans(torch.from_numpy(xt), a.nlist, niter=10).numpy()
    lab = assign(xt, coarse)
    pq = kmeans_subspaces(torch.from_numpy(xt - coarse[lab]), M, 256, niter=10).numpy()     # (M, 256, dsub)
    xb = tf(gen.chunk(SEED_BASE, 0, a.nb).numpy())
    ln = assign(xb, coarse)
    codes = encode(xb - coarse[ln], pq)
    order = np.argsort(ln, kind="stable")
    codes, lsorted = codes[order], ln[order]
    off = np.zeros(a.nlist + 1, np.int64)
    off[1:] = np.cumsum(np.bincount(ln, minlength=a.nlist))
    xq = tf(gen.chunk(SEED_QUERY, 0, a.nq).numpy())
    cn = (coarse * coarse).sum(1)
    probes = np.argsort(cn[None, :] - 2.0 * xq @ coarse.T, axis=1, kind="stable")[:, :a.nprobe]
    pqmax = np.sqrt((pq ** 2).sum(2)).max(1)
    cols = np.arange(M)[None, :]
    mu = coarse.mean(0)

    # per-vector term, float64, per-list grid
    pn2 = (pq.astype(np.float64) ** 2).sum(2)                                       # (M, 256)
    SB = np.empty(codes.shape[0], np.float64)
    for l in range(a.nlist):
        cc = codes[off[l]:off[l + 1]]
        cp = np.einsum("mkj,mj->mk", pq.astype(np.float64), (coarse[l] - mu).astype(np.float64).reshape(M, dsub))
        SB[off[l]:off[l + 1]] = (pn2 + 2.0 * cp)[cols, cc].sum(1)

    def quantise_sb(bits):
        out = np.empty_like(SB)
        for l in range(a.nlist):
            s = SB[off[l]:off[l + 1]]
            if s.shape[0] == 0:
                continue
            lo, hi = s.min(), s.max()
            step = max((hi - lo) / ((1 << bits) - 1), 1e-30)
            out[off[l]:off[l + 1]] = lo + np.floor((s - lo) / step) * step
        return out

    def run(mode, bits, sb_bits=16, centre=True):
        qmax = (1 << bits) - 1
        SBq = quantise_sb(sb_bits) if mode == "qlut" else None
        m0 = mu if centre else np.zeros_like(mu)
        surv = tot = 0
        for q in range(a.nq):
            best, thr = np.empty(0, np.float32), np.inf
            if mode == "qlut":
                A = -2.0 * np.einsum("mkj,mj->mk", pq.astype(np.float64), (xq[q] - m0).astype(np.float64).reshape(M, dsub))
                if not centre:
                    pass
                amin = A.min(1, keepdims=True)
                sq = qmax / ((A - amin).max() * 1.0001)
                UA = np.minimum(np.floor((A - amin) * sq), qmax).astype(np.int64)
                asum = float(amin.sum())
            for l in probes[q]:
                cc = codes[off[l]:off[l + 1]]
                if cc.shape[0] == 0:
                    continue
                r = (xq[q] - coarse[l]).reshape(M, 1, dsub)
                T = ((r - pq) ** 2).sum(2)                                            # (M, 256)
                ex = T[cols, cc].sum(1)
                if mode == "pair":
                    s = qmax / ((np.sqrt((r * r).sum(2))[:, 0] + pqmax) ** 2).max()
                    U = np.minimum(np.floor(T * s), qmax).astype(np.int64)
                    LB = U[cols, cc].sum(1)
                else:
                    LB = UA[cols, cc].sum(1)
                    dis0 = float((r.astype(np.float64) ** 2).sum())
                    sbl = SBq[off[l]:off[l + 1]]
                    if not centre:   # SB was built with mu: rebuild the uncentred per-vector term
                        cp = np.einsum("mkj,mj->mk", pq.astype(np.float64), coarse[l].astype(np.float64).reshape(M, dsub))
                        sbl = (pn2 + 2.0 * cp)[cols, cc].sum(1)
                for b0 in range(0, cc.shape[0], 256):
                    if not np.isfinite(thr):
                        passed = np.ones(min(256, cc.shape[0] - b0), bool)
                    elif mode == "pair":
                        passed = LB[b0:b0 + 256] <= np.floor(thr * s * 1.000004) + 1
                    else:
                        t = np.floor(sq * (thr * 1.000001 - dis0 - asum - sbl[b0:b0 + 256])) + 1
                        passed = LB[b0:b0 + 256] <= t
                    surv += int(passed.sum())
                    tot += passed.shape[0]
                    best = np.sort(np.concatenate([best, ex[b0:b0 + 256][passed]]))[:a.k]
                    if best.shape[0] == a.k:
                        thr = float(best[-1])
        return 100.0 * surv / tot, tot

    rows = []
    for name, kw in [("per-pair table, 11 bits (round 1)", dict(mode="pair", bits=11)),
                     ("per-query table 11 bits + SB 16 bits", dict(mode="qlut", bits=11, sb_bits=16)),
                     ("per-query table 10 bits + SB 16 bits", dict(mode="qlut", bits=10, sb_bits=16)),
                     ("per-query table 10 bits + SB 8 bits", dict(mode="qlut", bits=10, sb_bits=8)),
                     ("per-query table 8 bits + SB 16 bits", dict(mode="qlut", bits=8, sb_bits=16)),
                     ("per-query table 6 bits + SB 16 bits", dict(mode="qlut", bits=6, sb_bits=16)),
                     ("per-query table 11 bits + SB 16 bits, NOT centred", dict(mode="qlut", bits=11, sb_bits=16, centre=False))]:
        pct, tot = run(**kw)
        rows.append({"filter": name, "survivors_pct": round(pct, 3)})
        print(json.dumps(rows[-1]), file=sys.stderr, flush=True)
    print(json.dumps({"experiment": "per-query lower-bound filter (decomposed distance) vs per-pair filter: survivor rate",
                      "index": f"{a.nb} x {d}, IVF{a.nlist},PQ{M}, nprobe {a.nprobe}, k {a.k}, {a.nq} queries, clustered "
                               f"data x {a.scale} + {a.shift}", "threshold_refresh": "every 256 codes",
                      "codes_scanned": tot, "rows": rows}))


if __name__ == "__main__":
    main()
