"""Development probe (GPU): builds the C2-shaped index at --nb vectors once and times the search pipeline under
different environment settings (each setting gets a fresh handle on the same arrays).  Not a test, not a bench."""
import os, sys, json, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/chameleon-rag-acceleration_b200")
import bench
import b200ivfpq as faiss
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 12500000
settings = [dict(kv.split("=") for kv in a.split(",") if kv) for a in sys.argv[2:]] or [{}]
args = bench.parse_args(["--nb", str(nb), "--no-cpu-baseline", "--gt-queries", "0"])
cfg = (nb,) + bench.CONFIGS["c2"][1:]
dev = torch.device("cuda", 0)
index, xq, gt, info = bench.build_index(cfg, 0, 1, dev, None, args)
ref = None
for env in settings:
    for k_, v_ in env.items():
        os.environ[k_] = v_
    os.environ["B200_IVFPQ_QL_STATS"] = "1"
    ix = faiss.IndexIVFPQ(faiss.IndexFlatL2(128), 128, 8192, 16, 8)
    ix.set_codebooks(index.quantizer.xb_tensor(), index.pq.centroids_tensor())
    ix.set_lists(index._offsets, index._codes, index._ids)
    ix.nprobe = 32
    ix.set_stage_timing(True)
    for _ in range(4):
        D, I = ix.search(xq, 10)
    torch.cuda.synchronize()
    st = ix.stage_ms()
    if ref is None:
        ref = (D.clone(), I.clone())
    same = bool(torch.equal(D.view(torch.int32), ref[0].view(torch.int32)) and torch.equal(I, ref[1]))
    print(json.dumps({"env": env, "scan_ms": round(st["scan"], 3), "setup_ms": round(st["pair_setup"], 3), "filter_ms": ix.filter_ms(),
                      "stats": ix.filter_stats(reset=True), "same_as_first": same}), flush=True)
    for k_ in env:
        os.environ.pop(k_, None)
