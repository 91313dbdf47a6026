#!/usr/bin/env python
"""Multi-GPU parity check, run under torchrun (one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/check_multigpu.py

Every rank builds the same seeded index with the oracle (CPU), keeps its shard (add-order position % world),
and searches the whole batch through DistributedIndexIVFPQ (local CUDA search -> NCCL all-gather -> K5 merge).
Rank 0 compares the merged result with the oracle's search of the unsharded index.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "chameleon-rag-acceleration_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import b200ivfpq as faiss          # noqa: E402
from oracle import ivfpq_oracle as oracle   # noqa: E402  (checker only)
import _util                        # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for (d, nlist, M, n, nq, nprobe, k) in [(128, 64, 16, 40000, 200, 8, 10), (96, 32, 16, 9000, 50, 6, 100)]:
        a = _util.make_index_arrays(oracle, 5, d, nlist, M, n, id_scramble=False)
        xq = _util.make_queries(3, a, nq)
        full = faiss.IndexIVFPQ(faiss.IndexFlatL2(d), d, nlist, M, 8)
        full.set_codebooks(a["coarse"], a["pq"])
        full.set_lists(a["offsets"], a["codes"], a["ids"])
        for mode in ("vector", "list", "replica"):
            local_index = (faiss.shard_index(full, rank, world) if mode == "vector"
                           else faiss.shard_index_by_list(full, rank, world) if mode == "list" else full)
            local_index.nprobe = nprobe
            index = faiss.DistributedIndexIVFPQ(local_index, shard_mode=mode)
            D, I = index.search(torch.from_numpy(xq).cuda(), k)
            torch.cuda.synchronize()
            if rank == 0:
                Dr, Ir = oracle.C.search(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], nprobe, k)
                try:
                    _util.assert_same_modulo_ties(D.cpu().numpy(), I.cpu().numpy(), Dr, Ir, f"world {world} {mode}")
                    print(f"[check_multigpu] world={world} d={d} k={k} sharded by {mode}: merged result == oracle "
                          f"(modulo ties), shard sizes ~{local_index.ntotal}, "
                          f"merge = {'peer memory' if index.peer_merge else 'NCCL all-gather'}")
                    if mode == "replica":   # no merge at all: same kernels on the same lists as the single-GPU search
                        assert np.array_equal(I.cpu().numpy(), Ir), "replica ids differ from the oracle"
                except AssertionError as e:
                    ok = False
                    print("[check_multigpu] MISMATCH", e)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
