#!/usr/bin/env python
"""Batch-1 (and small-batch) latency of index.search(numpy) on a synthetic same-shape index."""
import argparse, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "chameleon-rag-acceleration_b200"))
from bench import CONFIGS
import b200ivfpq as faiss

ap = argparse.ArgumentParser(); ap.add_argument("--config", default="c2"); ap.add_argument("--nb", type=int, default=0)
args = ap.parse_args()
nb, d, nlist, M, nprobe, k, _ = CONFIGS[args.config]; nb = args.nb or nb
dev = torch.device("cuda", 0); g = torch.Generator(device=dev); g.manual_seed(1); rng = np.random.default_rng(1)
coarse = torch.rand((nlist, d), generator=g, device=dev); pq = torch.randn((M, 256, d // M), generator=g, device=dev) * 0.1
sizes = rng.multinomial(nb, np.full(nlist, 1.0 / nlist)); offsets = np.zeros(nlist + 1, np.int64); offsets[1:] = np.cumsum(sizes)
codes = torch.randint(0, 256, (nb, M), generator=g, device=dev, dtype=torch.uint8)
index = faiss.IndexIVFPQ(faiss.IndexFlatL2(d), d, nlist, M, 8); index.set_codebooks(coarse, pq); index.set_lists(offsets, codes, None)
index.nprobe = nprobe
out = {}
for bs in (1, 2, 4, 8, 16, 32, 64, 128):
    xq = np.random.default_rng(bs).random((256 + bs, d), dtype=np.float32)
    lat = []
    for i in range(120):
        q = xq[i:i + bs]
        t0 = time.perf_counter(); index.search(q, k); lat.append((time.perf_counter() - t0) * 1e3)
    lat = np.array(lat[20:]); out[bs] = {"p50_ms": float(np.median(lat)), "p95_ms": float(np.percentile(lat, 95))}
print(json.dumps({"config": args.config, "nb": nb, "nprobe": nprobe, "k": k, "graph": os.environ.get("B200_IVFPQ_GRAPH", "1"), "latency": out}, indent=1))
