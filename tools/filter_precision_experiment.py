#!/usr/bin/env python
"""CPU-only design experiment for the integer lower-bound filter of the four-query scan kernel (csrc/scan_quad.cuh):
how many codes survive the filter as a function of the precision of the quantised LUT entries?

A small clustered IVF-PQ index is built with numpy / torch on the CPU (same generator as bench.py); for every query the
probed lists are walked in scan order with the running k-th-best threshold refreshed every 256 codes (what the kernel
does per tile), and a code "survives" when its quantised lower bound does not exceed the threshold.  Two quantisers:
  plain   : u = floor(T * s), s = (2^bits - 1) / B, B = max_m (||r_m|| + max_c ||p_mc||)^2        (the shipped kernel)
  offset  : u = floor((T - o_m) * s), o_m = min_c T[m][c], s from max (T - o_m); threshold lowered by sum_m o_m
Nothing here touches the product path or the oracle; results go to stdout as JSON (profiles/r1_filter_precision_*.json).

    python tools/filter_precision_experiment.py [--nb 400000] [--nq 60]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "chameleon-rag-acceleration_b200"))
from b200ivfpq.datasets import SEED_BASE, SEED_QUERY, SEED_TRAIN, ClusteredGenerator  # noqa: E402
from b200ivfpq.kmeans import kmeans, kmeans_subspaces  # noqa: E402


def assign(x, c):
    out = np.empty(x.shape[0], np.int64)
    cn = (c * c).sum(1)
    for i0 in range(0, x.shape[0], 1 << 16):
        xb = x[i0:i0 + (1 << 16)]
        out[i0:i0 + xb.shape[0]] = (cn[None, :] - 2.0 * xb @ c.T).argmin(1)
    return out


def encode(res, pq):
    M, ksub, dsub = pq.shape
    codes = np.empty((res.shape[0], M), np.uint8)
    for m in range(M):
        sub = res[:, m * dsub:(m + 1) * dsub]
        pn = (pq[m] * pq[m]).sum(1)
        for i0 in range(0, sub.shape[0], 1 << 16):
            sb = sub[i0:i0 + (1 << 16)]
            codes[i0:i0 + sb.shape[0], m] = (pn[None, :] - 2.0 * sb @ pq[m].T).argmin(1)
    return codes


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nb", type=int, default=400000)
    ap.add_argument("--nq", type=int, default=60)
    ap.add_argument("--nlist", type=int, default=128)
    ap.add_argument("--nprobe", type=int, default=16)
    ap.add_argument("--k", type=int, default=10)
    a = ap.parse_args()
    d, M = 128, 16
    dsub = d // M
    gen = ClusteredGenerator(d, ncentres=a.nlist // 2, sigma=0.1, device=torch.device("cpu"), seed=7, latent_dim=12,
                             sigma_iso=0.002)
    xt = gen.chunk(SEED_TRAIN, 0, 40000)
    coarse = kmeans(xt, a.nlist, niter=10).numpy()
    lab = assign(xt.numpy(), coarse)
    pq = kmeans_subspaces(torch.from_numpy(xt.numpy() - coarse[lab]), M, 256, niter=10).numpy()     # (M, 256, dsub)
    xb = gen.chunk(SEED_BASE, 0, a.nb).numpy()
    ln = assign(xb, coarse)
    codes = encode(xb - coarse[ln], pq)
    order = np.argsort(ln, kind="stable")
    codes = codes[order]
    off = np.zeros(a.nlist + 1, np.int64)
    off[1:] = np.cumsum(np.bincount(ln, minlength=a.nlist))
    xq = gen.chunk(SEED_QUERY, 0, a.nq).numpy()
    cn = (coarse * coarse).sum(1)
    probes = np.argsort(cn[None, :] - 2.0 * xq @ coarse.T, axis=1, kind="stable")[:, :a.nprobe]
    pqmax = np.sqrt((pq ** 2).sum(2)).max(1)
    cols = np.arange(M)[None, :]

    def run(bits, offset):
        qmax = (1 << bits) - 1
        surv = tot = 0
        for q in range(a.nq):
            best, thr = np.empty(0, np.float32), np.inf
            for l in probes[q]:
                cc = codes[off[l]:off[l + 1]]
                if cc.shape[0] == 0:
                    continue
                r = (xq[q] - coarse[l]).reshape(M, 1, dsub)
                T = ((r - pq) ** 2).sum(2)                                            # (M, 256)
                if offset:
                    om = T.min(1, keepdims=True)
                    s = qmax / ((T - om).max() * 1.0001)
                    U = np.minimum(np.floor((T - om) * s), qmax).astype(np.int64)
                    osum = float(om.sum()) * 0.99999
                else:
                    s = qmax / ((np.sqrt((r * r).sum(2))[:, 0] + pqmax) ** 2).max()
                    U = np.minimum(np.floor(T * s), qmax).astype(np.int64)
                    osum = 0.0
                LB, ex = U[cols, cc].sum(1), T[cols, cc].sum(1)
                for b0 in range(0, cc.shape[0], 256):
                    t_int = np.floor(max(thr - osum, 0.0) * s * 1.000004) + 1 if np.isfinite(thr) else 1 << 40
                    passed = LB[b0:b0 + 256] <= t_int
                    surv += int(passed.sum())
                    tot += passed.shape[0]
                    best = np.sort(np.concatenate([best, ex[b0:b0 + 256][passed]]))[:a.k]
                    if best.shape[0] == a.k:
                        thr = float(best[-1])
        return surv, tot

    rows = []
    for bits in (11, 10, 8, 6, 5, 4, 3, 2):
        s1, t = run(bits, False)
        s2, _ = run(bits, True)
        rows.append({"bits": bits, "survivors_pct_plain": round(100.0 * s1 / t, 3),
                     "survivors_pct_offset": round(100.0 * s2 / t, 3)})
        print(json.dumps(rows[-1]), file=sys.stderr, flush=True)
    print(json.dumps({"experiment": "lower-bound filter: survivor rate vs LUT entry precision (CPU simulation)",
                      "index": f"{a.nb} x {d}, IVF{a.nlist},PQ{M}, nprobe {a.nprobe}, k {a.k}, {a.nq} queries, clustered data",
                      "threshold_refresh": "every 256 codes", "codes_scanned": t, "rows": rows}))


if __name__ == "__main__":
    main()
