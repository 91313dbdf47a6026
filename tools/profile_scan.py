#!/usr/bin/env python
"""Short driver for ncu / timing experiments on the scan kernel: synthetic same-shape index (uniform random
codes, multinomial list sizes -- no training or encoding), a few searches, per-stage CUDA-event times.

    python tools/profile_scan.py --config c2 --iters 3 [--nb N] [--nq N] [--k K] [--nprobe P]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "chameleon-rag-acceleration_b200"))
from bench import CONFIGS  # noqa: E402

import b200ivfpq as faiss  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--nb", type=int, default=0)
    ap.add_argument("--nq", type=int, default=0)
    ap.add_argument("--k", type=int, default=0)
    ap.add_argument("--nprobe", type=int, default=0)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--skew", type=float, default=0.0, help="dirichlet concentration for list sizes (0 = balanced)")
    args = ap.parse_args()
    nb, d, nlist, M, nprobe, k, nq = CONFIGS[args.config]
    nb, nq, k, nprobe = args.nb or nb, args.nq or nq, args.k or k, args.nprobe or nprobe
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    rng = np.random.default_rng(1)
    coarse = torch.rand((nlist, d), generator=g, device=dev)
    pq = torch.randn((M, 256, d // M), generator=g, device=dev) * 0.1
    p = np.full(nlist, 1.0 / nlist) if args.skew <= 0 else rng.dirichlet(np.full(nlist, args.skew))
    sizes = rng.multinomial(nb, p)
    offsets = np.zeros(nlist + 1, np.int64)
    offsets[1:] = np.cumsum(sizes)
    codes = torch.randint(0, 256, (nb, M), generator=g, device=dev, dtype=torch.uint8)
    index = faiss.IndexIVFPQ(faiss.IndexFlatL2(d), d, nlist, M, 8)
    index.set_codebooks(coarse, pq)
    index.set_lists(offsets, codes, None)
    index.nprobe = nprobe
    index.set_stage_timing(True)
    xq = torch.rand((nq, d), generator=g, device=dev)
    out = []
    for it in range(args.iters):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        D, I = index.search(xq, k)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        st = index.stage_ms()
        stats = index.last_scan_stats()
        out.append({"wall_ms": wall, **st, "scan_GBps": stats["bytes"] / st["scan"] / 1e6,
                    "bytes_per_query_MB": stats["bytes"] / nq / 1e6})
    print(json.dumps({"config": args.config, "nb": nb, "nq": nq, "k": k, "nprobe": nprobe,
                      "variant": os.environ.get("B200_IVFPQ_SCAN", "auto"), "iters": out}, indent=1))


if __name__ == "__main__":
    main()
