"""The N > 1 path on CPU: world_size 2 over gloo.  The exchange logic of DistributedIndexIVFPQ (pack, one
all-gather, unpack, merge) runs for real; the per-shard search and the merge kernel -- CUDA on the product
path -- are replaced by the oracle through the injection points, which is the only way to exercise the
collective plumbing without a GPU."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import _util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "chameleon-rag-acceleration_b200"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import torch.distributed as dist
    from oracle import ivfpq_oracle as oracle
    from b200ivfpq.shards import DistributedIndexIVFPQ
    import _util as U
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a = U.make_index_arrays(oracle, 17, 32, 16, 8, 3000, id_scramble=False)
        xq = U.make_queries(6, a, 24)
        nprobe, k = 5, 10
        sizes = np.diff(a["offsets"])
        list_no = np.repeat(np.arange(16), sizes)
        keep = (a["ids"] % world) == rank                     # shard by add-order position
        off = np.zeros(17, np.int64)
        off[1:] = np.cumsum(np.bincount(list_no[keep], minlength=16))
        codes, ids = a["codes"][keep], a["ids"][keep]

        def local_search(x, kk):
            D, I = oracle.C.search(x.numpy(), a["coarse"], a["pq"], off, codes, ids, nprobe, kk)
            return torch.from_numpy(D), torch.from_numpy(I)

        def merge(Ds, Is):
            D, I = oracle.C.merge_shards(Ds.numpy(), Is.numpy())
            return torch.from_numpy(D), torch.from_numpy(I)

        class _Local:
            nprobe = 5
            d = 32

        index = DistributedIndexIVFPQ(_Local(), merge_fn=merge, local_search_fn=local_search)
        assert index.world == world and index.rank == rank
        D, I = index.search(torch.from_numpy(xq), k)
        # replicated index, batch sliced by query (co.shard = False): unequal slices (23 queries) and a batch smaller
        # than the world (1 query: rank 1 only joins the exchange)
        def full_search(x, kk):
            D_, I_ = oracle.C.search(x.numpy(), a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], nprobe, kk)
            return torch.from_numpy(D_), torch.from_numpy(I_)

        rep = DistributedIndexIVFPQ(_Local(), local_search_fn=full_search, shard_mode="replica")
        Dr, Ir = rep.search(torch.from_numpy(xq[:23]), k)
        D1, I1 = rep.search(torch.from_numpy(xq[:1]), k)

        # the reference's entry point: index_cpu_to_gpu_multiple(vres, vdev, index, co) with co.shard = False is
        # Faiss's IndexReplicas -> inside a process group, a replica that answers its slice of the batch
        import b200ivfpq as faiss

        class _Whole:
            nprobe = 5
            d = 32
            search = staticmethod(full_search)

        co = faiss.GpuMultipleClonerOptions()
        assert co.shard is False
        cloned = faiss.index_cpu_to_gpu_multiple(None, None, _Whole(), co)
        assert isinstance(cloned, DistributedIndexIVFPQ) and cloned.shard_mode == "replica" and not cloned.peer_merge
        Dc, Ic = cloned.search(torch.from_numpy(xq[:23]), k)
        assert torch.equal(Dc, Dr) and torch.equal(Ic, Ir)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), D=D.numpy(), I=I.numpy(), Dr=Dr.numpy(), Ir=Ir.numpy(),
                 D1=D1.numpy(), I1=I1.numpy())
    finally:
        dist.destroy_process_group()


def _worker_rxs(rank, world, port, out_dir):
    """2 replicas x 2 shards (the reference's `-R 2` on 4 GPUs, bench_gpu_performance_OSDI.py:613-626)."""
    for p in (ROOT, os.path.join(ROOT, "chameleon-rag-acceleration_b200"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import torch.distributed as dist
    from oracle import ivfpq_oracle as oracle
    from b200ivfpq.shards import DistributedIndexIVFPQ, IndexReplicas, make_replica_groups, replica_layout
    import _util as U
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        R = 2
        S, rep, sh = replica_layout(world, rank, R)
        assert (S, rep, sh) == (2, rank // 2, rank % 2)
        shard_group, cross_group = make_replica_groups(R)
        a = U.make_index_arrays(oracle, 17, 32, 16, 8, 3000, id_scramble=False)
        xq = U.make_queries(6, a, 24)
        nprobe, k = 5, 10
        sizes = np.diff(a["offsets"])
        list_no = np.repeat(np.arange(16), sizes)
        keep = (a["ids"] % S) == sh                           # this replica's shard `sh` of S
        off = np.zeros(17, np.int64)
        off[1:] = np.cumsum(np.bincount(list_no[keep], minlength=16))
        codes, ids = a["codes"][keep], a["ids"][keep]
        seen = []

        def local_search(x, kk):
            seen.append(x.shape[0])
            D, I = oracle.C.search(x.numpy(), a["coarse"], a["pq"], off, codes, ids, nprobe, kk)
            return torch.from_numpy(D), torch.from_numpy(I)

        def merge(Ds, Is):
            D, I = oracle.C.merge_shards(Ds.numpy(), Is.numpy())
            return torch.from_numpy(D), torch.from_numpy(I)

        class _Local:
            nprobe = 5
            d = 32

        inner = DistributedIndexIVFPQ(_Local(), group=shard_group, merge_fn=merge, local_search_fn=local_search)
        assert inner.world == S and inner.rank == sh and not inner.peer_merge
        index = IndexReplicas(inner, R, rep, cross_group)
        D, I = index.search(torch.from_numpy(xq[:23]), k)
        assert seen == [11 if rep == 0 else 12]               # each replica searched only its slice
        D1, I1 = index.search(torch.from_numpy(xq[:1]), k)    # batch 1: replica 1 (the last slice) answers
        assert seen == ([11] if rep == 0 else [12, 1])
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), D=D.numpy(), I=I.numpy(), D1=D1.numpy(), I1=I1.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world4_two_replicas_of_two_shards(oracle, tmp_path):
    world = 4
    port = _free_port()
    mp.spawn(_worker_rxs, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    a = _util.make_index_arrays(oracle, 17, 32, 16, 8, 3000, id_scramble=False)
    xq = _util.make_queries(6, a, 24)
    D, I = oracle.C.search(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], 5, 10)
    outs = [np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(world)]
    for r in range(world):
        _util.assert_bit_equal(outs[r]["D"], outs[0]["D"], f"rank {r} agrees with rank 0 on D")
        _util.assert_bit_equal(outs[r]["I"], outs[0]["I"], f"rank {r} agrees with rank 0 on I")
        _util.assert_same_modulo_ties(outs[r]["D"], outs[r]["I"], D[:23], I[:23], f"2x2 layout vs single index, rank {r}")
        _util.assert_same_modulo_ties(outs[r]["D1"], outs[r]["I1"], D[:1], I[:1], f"2x2 layout batch 1, rank {r}")


def test_replica_layout_rules():
    sys.path.insert(0, os.path.join(ROOT, "chameleon-rag-acceleration_b200"))
    from b200ivfpq.shards import replica_layout
    assert replica_layout(8, 5, 1) == (8, 0, 5)
    assert replica_layout(8, 5, 2) == (4, 1, 1)
    assert replica_layout(8, 5, 8) == (1, 5, 0)
    with pytest.raises(ValueError):
        replica_layout(8, 0, 3)


@pytest.mark.timeout(300)
def test_world2_gloo_exchange_and_merge(oracle, tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    a = _util.make_index_arrays(oracle, 17, 32, 16, 8, 3000, id_scramble=False)
    xq = _util.make_queries(6, a, 24)
    D, I = oracle.C.search(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], 5, 10)
    outs = [np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(world)]
    _util.assert_bit_equal(outs[0]["D"], outs[1]["D"], "ranks agree on D")
    _util.assert_bit_equal(outs[0]["I"], outs[1]["I"], "ranks agree on I")
    _util.assert_same_modulo_ties(outs[0]["D"], outs[0]["I"], D, I, "sharded vs single index")
    for r in range(world):
        _util.assert_bit_equal(outs[r]["Dr"], D[:23], f"replica D on rank {r}")
        _util.assert_bit_equal(outs[r]["Ir"], I[:23], f"replica I on rank {r}")
        _util.assert_bit_equal(outs[r]["D1"], D[:1], f"replica batch-1 D on rank {r}")
        _util.assert_bit_equal(outs[r]["I1"], I[:1], f"replica batch-1 I on rank {r}")
