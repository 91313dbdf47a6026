import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/chameleon-rag-acceleration_b200")
os.environ["B200_IVFPQ_QL_STATS"]="1"
import bench
args = bench.parse_args(["--nb","12500000","--no-cpu-baseline","--gt-queries","0"])
cfg=(12500000,)+bench.CONFIGS["c2"][1:]
dev=torch.device("cuda",0)
index,xq,gt,info=bench.build_index(cfg,0,1,dev,None,args)
import b200ivfpq as faiss
a=None
for capq in (1024, 4096, 16384, 65536):
    os.environ["B200_IVFPQ_STREAM_CAPQ"]=str(capq)
    ix = faiss.IndexIVFPQ(faiss.IndexFlatL2(128), 128, 8192, 16, 8)
    ix.set_codebooks(index.quantizer.xb_tensor(), index.pq.centroids_tensor())
    ix.set_lists(index._offsets, index._codes, index._ids)
    ix.nprobe=32
    ix.set_stage_timing(True)
    for nq in (10000,):
        for _ in range(3): D,I=ix.search(xq[:nq],10)
        torch.cuda.synchronize()
        print(capq, nq, ix.filter_stats(reset=True), ix.stage_ms(), flush=True)

