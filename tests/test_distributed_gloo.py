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


class _OracleQuantizer:
    """index.quantizer stand-in on CPU tensors: the oracle's coarse stage."""

    def __init__(self, oracle, cent):
        self.oracle, self.cent, self.calls = oracle, cent, []

    def search(self, x, k):
        self.calls.append(int(x.shape[0]))
        D, I = self.oracle.C.coarse(x.numpy(), self.cent, k)
        return torch.from_numpy(D), torch.from_numpy(I)


class _OracleShard:
    """Rank-local IndexIVFPQ stand-in: quantizer.search / search / search_preassigned answered by the oracle on this
    rank's shard.  Lets DistributedIndexIVFPQ._local_search itself (sliced coarse stage, probe exchange, by-list probe
    masking) run under gloo; only the merge kernel is injected."""

    def __init__(self, oracle, a, off, codes, ids, nprobe):
        self.oracle, self.a, self.off, self.codes, self.ids = oracle, a, off, codes, ids
        self.nprobe, self.d, self.nlist = nprobe, a["coarse"].shape[1], a["coarse"].shape[0]
        self.quantizer = _OracleQuantizer(oracle, a["coarse"])
        self.preassigned = []

    def search(self, x, k):
        D, I = self.oracle.C.search(x.numpy(), self.a["coarse"], self.a["pq"], self.off, self.codes, self.ids,
                                    self.nprobe, k)
        return torch.from_numpy(D), torch.from_numpy(I)

    def search_preassigned(self, x, k, probes):
        self.preassigned.append(probes.clone())
        D, I = self.oracle.C.search_preassigned(x.numpy(), self.a["coarse"], self.a["pq"], self.off, self.codes,
                                                self.ids, probes.numpy(), k)
        return torch.from_numpy(D), torch.from_numpy(I)


class _OracleShardSplit(_OracleShard):
    """The same stand-in with the two-half search of the threshold exchange (IndexIVFPQ.search_preassigned_begin /
    _finish): begin answers the k-th best LOCAL distance of the queries in its slice (+inf bits elsewhere), finish drops
    every local result above the exchanged threshold -- what the CUDA filter does with it."""

    def search_preassigned_begin(self, x, k, probes, lo, hi):
        self._x, self._k, self._probes = x, k, probes
        thr = torch.full((x.shape[0],), 0x7f800000, dtype=torch.int32)
        D, _ = self.oracle.C.search_preassigned(x[lo:hi].numpy(), self.a["coarse"], self.a["pq"], self.off, self.codes,
                                                self.ids, probes[lo:hi].numpy(), k)
        kth = torch.from_numpy(np.ascontiguousarray(D[:, k - 1])).view(torch.int32)
        full = torch.from_numpy(D[:, k - 1] < 3e38)                    # fewer than k local results: no threshold
        thr[lo:hi] = torch.where(full, kth, thr[lo:hi])
        self.begun = (lo, hi)
        return thr

    def search_preassigned_finish(self, thr, nq, k, out=None):
        self.exchanged = thr.clone()
        D, I = self.oracle.C.search_preassigned(self._x.numpy(), self.a["coarse"], self.a["pq"], self.off, self.codes,
                                                self.ids, self._probes.numpy(), k)
        t = thr.view(torch.float32).numpy()[:, None]
        drop = D > t
        D[drop], I[drop] = np.finfo(np.float32).max, -1
        return torch.from_numpy(D), torch.from_numpy(I)


def _worker_threshold_exchange(rank, world, port, out_dir):
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
        a = U.make_index_arrays(oracle, 23, 32, 16, 8, 4000, id_scramble=False)
        xq = torch.from_numpy(U.make_queries(9, a, 40))
        nprobe, k, nlist = 6, 10, 16
        sizes = np.diff(a["offsets"])
        list_no = np.repeat(np.arange(nlist), sizes)
        keep = (a["ids"] % world) == rank
        off = np.zeros(nlist + 1, np.int64)
        off[1:] = np.cumsum(np.bincount(list_no[keep], minlength=nlist))
        local = _OracleShardSplit(oracle, a, off, a["codes"][keep], a["ids"][keep], nprobe)

        def merge(Ds, Is):
            D, I = oracle.C.merge_shards(Ds.numpy(), Is.numpy())
            return torch.from_numpy(D), torch.from_numpy(I)

        index = DistributedIndexIVFPQ(local, merge_fn=merge)
        index.exchange_thresholds = True            # (off by default for injected merges: this test IS the exchange)
        D, I = index.search(xq, k)
        lo, hi = local.begun
        assert (lo, hi) == ((40 * rank) // world, (40 * (rank + 1)) // world)
        ex = local.exchanged.view(torch.float32)
        assert bool(torch.isfinite(ex).all()), "every query got a threshold from some rank"
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), D=D.numpy(), I=I.numpy(), thr=ex.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 4])
def test_threshold_exchange_keeps_the_result(oracle, tmp_path, world):
    """Multi-GPU threshold exchange (shards.py): every rank bootstraps the thresholds of its slice of the queries on its
    own shard, an all-reduce MIN hands all of them to every rank, every shard drops what lies above -- and the merged
    result is still the unsharded oracle's, because any shard's k-th best distance bounds the global k-th distance."""
    port = _free_port()
    mp.spawn(_worker_threshold_exchange, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    a = _util.make_index_arrays(oracle, 23, 32, 16, 8, 4000, id_scramble=False)
    xq = _util.make_queries(9, a, 40)
    D, I = oracle.C.search(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], 6, 10)
    outs = [np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(world)]
    for r in range(world):
        _util.assert_same_modulo_ties(outs[r]["D"], outs[r]["I"], D, I, f"rank {r}")
        assert np.array_equal(outs[r]["thr"], outs[0]["thr"]), "all ranks hold the same thresholds"
        assert (outs[r]["thr"] >= D[:, -1]).all(), "a threshold below the true k-th distance would lose results"


def _worker_local_search(rank, world, port, out_dir):
    """vector and list layouts through the REAL _local_search (no local_search_fn injection)."""
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
        xq = torch.from_numpy(U.make_queries(6, a, 24))
        nprobe, k, nlist = 5, 10, 16
        sizes = np.diff(a["offsets"])
        list_no = np.repeat(np.arange(nlist), sizes)

        def merge(Ds, Is):
            D, I = oracle.C.merge_shards(Ds.numpy(), Is.numpy())
            return torch.from_numpy(D), torch.from_numpy(I)

        out = {}
        for mode in ("vector", "list"):
            keep = ((a["ids"] % world) == rank) if mode == "vector" else ((list_no % world) == rank)
            off = np.zeros(nlist + 1, np.int64)
            off[1:] = np.cumsum(np.bincount(list_no[keep], minlength=nlist))
            local = _OracleShard(oracle, a, off, a["codes"][keep], a["ids"][keep], nprobe)
            index = DistributedIndexIVFPQ(local, merge_fn=merge, shard_mode=mode)
            assert not index.peer_merge
            D, I = index.search(xq, k)                       # 24 queries >= 8 * world: the coarse stage is sliced
            assert local.quantizer.calls == [12], "each rank ranks the centroids for its half of the batch only"
            probes = local.preassigned[0]
            assert probes.shape == (24, nprobe)
            if mode == "list":
                own = probes >= 0
                assert bool(((probes[own] % world) == rank).all()) and bool((~own).any()), "foreign lists masked"
            D3, I3 = index.search(xq[:3], k)                 # 3 queries < 8 * world: no slicing
            if mode == "vector":
                assert local.quantizer.calls == [12]         # plain local search, quantizer not called separately
            else:
                assert local.quantizer.calls == [12, 3]      # by-list always needs the probes to mask them
            out[mode] = (D.numpy(), I.numpy(), D3.numpy(), I3.numpy())
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **{f"{m}_{n}": v for m, t in out.items()
                                                              for n, v in zip(("D", "I", "D3", "I3"), t)})
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world2_sliced_coarse_and_by_list_masking(oracle, tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker_local_search, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    a = _util.make_index_arrays(oracle, 17, 32, 16, 8, 3000, id_scramble=False)
    xq = _util.make_queries(6, a, 24)
    D, I = oracle.C.search(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], 5, 10)
    outs = [np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(world)]
    for mode in ("vector", "list"):
        for r in range(world):
            _util.assert_same_modulo_ties(outs[r][f"{mode}_D"], outs[r][f"{mode}_I"], D, I, f"{mode} rank {r}")
            _util.assert_same_modulo_ties(outs[r][f"{mode}_D3"], outs[r][f"{mode}_I3"], D[:3], I[:3],
                                          f"{mode} small batch rank {r}")


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
