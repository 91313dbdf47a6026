#!/usr/bin/env python
"""BASELINE.json config 5: batch x nprobe sweep on the C5 shape (100M x 128, IVF8192,PQ32x8) to map where the scan is
bound by code bytes (algorithmic GB/s flat), by per-(query, list) work (LUT build, top-k: small lists per query) or by
latency (small batches).  Synthetic same-shape index: uniform random codes, multinomial list sizes (scan-bandwidth
only, no recall; SURVEY.md section 8d "scan-only alternative").

    python tools/sweep_c5.py [--config c5] [--nb N] > gpurun_out/sweep_c5.json
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
    ap.add_argument("--config", default="c5")
    ap.add_argument("--nb", type=int, default=0)
    ap.add_argument("--batches", default="1,4,16,64,256,1024,4096")
    ap.add_argument("--nprobes", default="1,4,16,64,256")
    args = ap.parse_args()
    nb, d, nlist, M, _, k, _ = CONFIGS[args.config]
    nb = args.nb or nb
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    rng = np.random.default_rng(1)
    coarse = torch.rand((nlist, d), generator=g, device=dev)
    pq = torch.randn((M, 256, d // M), generator=g, device=dev) * 0.1
    sizes = rng.multinomial(nb, np.full(nlist, 1.0 / nlist))
    offsets = np.zeros(nlist + 1, np.int64)
    offsets[1:] = np.cumsum(sizes)
    codes = torch.randint(0, 256, (nb, M), generator=g, device=dev, dtype=torch.uint8)
    index = faiss.IndexIVFPQ(faiss.IndexFlatL2(d), d, nlist, M, 8)
    index.set_codebooks(coarse, pq)
    index.set_lists(offsets, codes, None)
    index.set_stage_timing(True)
    rows = []
    for bs in [int(x) for x in args.batches.split(",")]:
        xq = torch.rand((bs, d), generator=g, device=dev)
        for nprobe in [int(x) for x in args.nprobes.split(",")]:
            index.nprobe = nprobe
            for _ in range(2):
                index.search(xq, k)
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 5 if bs * nprobe <= 65536 else 2
            ev0.record()
            for _ in range(iters):
                index.search(xq, k)
            ev1.record()
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1) / iters
            st = index.stage_ms()
            stats = index.last_scan_stats()
            rows.append({"batch": bs, "nprobe": nprobe, "ms": round(ms, 4), "qps": round(bs / ms * 1e3, 1),
                         "scan_ms": round(st["scan"], 4), "coarse_ms": round(st["coarse_dist"] + st["coarse_select"], 4),
                         "scan_GBps": round(stats["bytes"] / st["scan"] / 1e6, 1),
                         "queries_per_list": round(bs * nprobe / nlist, 2)})
    print(json.dumps({"config": args.config, "nb": nb, "d": d, "nlist": nlist, "M": M, "k": k,
                      "data": "synthetic same-shape index (uniform random codes)", "device_resident_queries": True,
                      "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
