#!/usr/bin/env python
"""Per-shard scan time of the by-list split, measured one shard after the other on ONE GPU.

Builds bench.py's index (same generator, training and encoding), then for every rank r of a world of W keeps the lists
l % W == r (shard_index_by_list), masks the other probes and times search_preassigned.  A projection aid for sizing and
for finding imbalance -- the multi-GPU numbers that count are bench.py's under torchrun.

    python tools/list_shard_probe.py --config c2 --worlds 1,2,8 [--scan N]
"""
import argparse
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "chameleon-rag-acceleration_b200"))
import bench  # noqa: E402

import b200ivfpq as faiss  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--nb", type=int, default=0)
    ap.add_argument("--worlds", default="1,2,8")
    ap.add_argument("--iters", type=int, default=4)
    a = ap.parse_args()
    cfg = bench.CONFIGS[a.config]
    if a.nb:
        cfg = (a.nb,) + cfg[1:]
    nb, d, nlist, M, nprobe, k, nq = cfg
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    bargs = bench.parse_args([]) if hasattr(bench, "parse_args") else None
    if bargs is None:
        raise SystemExit("bench.parse_args missing")
    bargs.gt_queries = 0
    full, xq, _, _ = bench.build_index(cfg, 0, 1, dev, None, bargs)
    _, probes = full.quantizer.search(xq, nprobe)
    sizes_full = np.diff(full._finalized_offsets())
    out = []
    for W in [int(w) for w in a.worlds.split(",")]:
        for r in range(W):
            sub = faiss.shard_index_by_list(full, r, W) if W > 1 else full
            sub.nprobe = nprobe
            sub.set_stage_timing(True)
            masked = torch.where(probes % W == r, probes, torch.full_like(probes, -1))
            scan = []
            for _ in range(a.iters):
                sub.search_preassigned(xq, k, masked)
                torch.cuda.synchronize()
                scan.append(sub.stage_ms()["scan"])
            st = sub.last_scan_stats()
            mine = sizes_full[np.arange(nlist) % W == r]
            row = {"world": W, "rank": r, "scan_ms": round(float(np.median(scan[1:])), 3),
                   "gbytes": round(st["bytes"] / 1e9, 2), "gbs": round(st["bytes"] / 1e6 / float(np.median(scan[1:])), 0),
                   "ntotal": int(sub.ntotal), "valid_pairs": int((masked >= 0).sum()),
                   "max_list": int(mine.max()), "p99_list": int(np.percentile(mine, 99))}
            print(json.dumps(row), file=sys.stderr, flush=True)
            out.append(row)
            if W > 1:
                del sub
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "list_shard_probe.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
