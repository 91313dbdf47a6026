"""Multi-GPU search: inverted lists sharded by vector, per-shard top-k merged after an all-gather (the default, and
what the reference runs); optionally whole lists per GPU, or a replicated index with the batch sliced by query.

Reference behaviour: faiss.index_cpu_to_gpu_multiple(vres, vdev, index, co) with co.shard = True
(bench_gpu_performance_OSDI.py:586-604), i.e. Faiss IndexShards with the modulo split the reference logs
("IndexShards shard 0 select modulo 3 = 0", Faiss_experiments/gpu_recall:3-7); the merge is concatenate +
argsort + take k (bench_multi_cpu_performance_OSDI.py:203-219).

B200 design: one process per GPU (torchrun).  Every rank holds the same codebooks and
the entries whose position in add order is congruent to its rank; it searches the WHOLE query batch against
its shard, then one NCCL all-gather of the packed per-shard (D, I) over NVLink and the K5 merge kernel on
every rank.  No other collective exists on the path.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch

from . import _lib
from .index import IndexIVFPQ


def shard_positions(ntotal: int, rank: int, world: int) -> np.ndarray:
    """Positions (in add order) owned by `rank`: i % world == rank."""
    return np.arange(rank, ntotal, world, dtype=np.int64)


def shard_index(index: IndexIVFPQ, rank: int, world: int) -> IndexIVFPQ:
    """Build rank's shard of a populated index (same codebooks, entries with add-order position
    % world == rank).  Add order is recovered from the stored ids when they are the default sequential
    ids; for user ids the split is by position inside each list, which balances equally well."""
    index._finalize_lists()
    sub = IndexIVFPQ(None, index.d, index.nlist, index.pq.M, index.pq.nbits)
    sub.set_codebooks(index.quantizer.xb_tensor(), index.pq.centroids_tensor())
    sub.nprobe = index.nprobe
    off = index._offsets
    ntotal = int(off[-1])
    if ntotal == 0:
        return sub
    dev = index._codes.device
    ids = index._ids
    seq = bool(torch.equal(torch.sort(ids)[0], torch.arange(ntotal, device=dev)))
    key = ids if seq else torch.arange(ntotal, device=dev)
    keep = (key % world) == rank
    sizes = torch.from_numpy(np.diff(off)).to(dev)
    list_no = torch.repeat_interleave(torch.arange(index.nlist, device=dev), sizes)
    counts = torch.bincount(list_no[keep], minlength=index.nlist).cpu().numpy()
    new_off = np.zeros(index.nlist + 1, np.int64)
    new_off[1:] = np.cumsum(counts)
    sub.set_lists(new_off, index._codes[keep].contiguous(), ids[keep].contiguous())
    return sub


def shard_index_by_list(index: IndexIVFPQ, rank: int, world: int) -> IndexIVFPQ:
    """Rank's shard when WHOLE LISTS are distributed: list l lives on rank l % world (all nlist lists exist on every
    rank, the others are empty).  Not what the reference does (Faiss IndexShards splits by vector), but the layout that
    removes the per-(query, list) work every by-vector shard repeats: a (query, list) pair is then scanned -- and its
    LUT built -- on exactly one GPU.  Use with DistributedIndexIVFPQ(..., shard_mode="list")."""
    index._finalize_lists()
    sub = IndexIVFPQ(None, index.d, index.nlist, index.pq.M, index.pq.nbits)
    sub.set_codebooks(index.quantizer.xb_tensor(), index.pq.centroids_tensor())
    sub.nprobe = index.nprobe
    off = index._offsets
    if int(off[-1]) == 0:
        return sub
    dev = index._codes.device
    sizes = np.diff(off)
    mine = (np.arange(index.nlist) % world) == rank
    keep_rows = torch.repeat_interleave(torch.from_numpy(mine).to(dev), torch.from_numpy(sizes).to(dev))
    new_off = np.zeros(index.nlist + 1, np.int64)
    new_off[1:] = np.cumsum(np.where(mine, sizes, 0))
    sub.set_lists(new_off, index._codes[keep_rows].contiguous(), index._ids[keep_rows].contiguous())
    return sub


def merge_shards(Ds: torch.Tensor, Is: torch.Tensor):
    """K5 on the current device.  Ds, Is: (nshard, nq, k) CUDA tensors -> (D, I) (nq, k)."""
    lib = _lib.load()
    nshard, nq, k = Ds.shape
    Ds, Is = Ds.contiguous(), Is.contiguous()
    D = torch.empty((nq, k), dtype=torch.float32, device=Ds.device)
    I = torch.empty((nq, k), dtype=torch.int64, device=Ds.device)
    with torch.cuda.device(Ds.device):
        st = int(torch.cuda.current_stream(Ds.device).cuda_stream)
        _lib.check(lib.b200_ivfpq_merge_shards(nshard, nq, k, Ds.data_ptr(), Is.data_ptr(), D.data_ptr(), I.data_ptr(),
                                               st))
    return D, I


def pack_results(D: torch.Tensor, I: torch.Tensor) -> torch.Tensor:
    """(nq, k) f32 + (nq, k) i64 -> one (nq, k, 3) int32 buffer so that the exchange is a single all-gather."""
    out = torch.empty(D.shape + (3,), dtype=torch.int32, device=D.device)
    out[..., 0] = D.view(torch.int32)
    out[..., 1:] = I.view(torch.int32).reshape(I.shape + (2,))
    return out


def unpack_results(buf: torch.Tensor):
    D = buf[..., 0].contiguous().view(torch.float32)
    I = buf[..., 1:].contiguous().view(torch.int64).reshape(buf.shape[:-1])
    return D, I


def all_gather_rows(dist, group, world: int, mine: torch.Tensor, counts):
    """all-gather of row blocks of unequal height (padded to the largest block): rank r of `group` contributes
    counts[r] rows; everyone gets the blocks stacked in rank order."""
    hmax = max(counts)
    pad = mine
    if mine.shape[0] < hmax:
        pad = torch.cat([mine, mine.new_zeros((hmax - mine.shape[0],) + mine.shape[1:])], 0)
    out = torch.empty((world * hmax,) + mine.shape[1:], dtype=mine.dtype, device=mine.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    out = out.view((world, hmax) + mine.shape[1:])
    return torch.cat([out[r, :counts[r]] for r in range(world)], 0)


def replica_layout(world: int, rank: int, replicas: int):
    """R replica groups of S = world / R consecutive ranks (the reference's `-R`: GPUs [ngpu*i/R, ngpu*(i+1)/R) form
    replica i, bench_gpu_performance_OSDI.py:613-626).  Returns (S, replica id, shard rank inside the replica)."""
    if replicas < 1 or world % replicas:
        raise ValueError(f"replicas = {replicas} must divide the world size {world}")
    S = world // replicas
    return S, rank // S, rank % S


def make_replica_groups(replicas: int):
    """Collective (every rank of the default group calls it): the process groups of the R x S layout.  Returns
    (shard_group, cross_group): my replica's S ranks (probe exchange + top-k merge) and the R ranks holding the same
    shard in the other replicas (exchange of the per-replica result slices)."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    S, rep, sh = replica_layout(world, rank, replicas)
    shard_group = cross_group = None
    for r in range(replicas):                       # new_group is collective over the default group: same order everywhere
        g = dist.new_group(list(range(r * S, (r + 1) * S)))
        if r == rep:
            shard_group = g
    for j in range(S):
        g = dist.new_group(list(range(j, world, S)))
        if j == sh:
            cross_group = g
    return shard_group, cross_group


class _FnIndex:
    """search(xq, k) from a plain function (the CPU tests' injection point)."""

    def __init__(self, fn):
        self.search = fn


class IndexReplicas:
    """Faiss IndexReplicas in SPMD form: R replicas of the same index, every batch sliced by query between them.

    `inner` answers this replica's slice -- an IndexIVFPQ holding the whole index (R = world), or a
    DistributedIndexIVFPQ over this replica's shard group (R replicas x S shards).  `cross_group` holds one rank of
    every replica (make_replica_groups); the slices are exchanged with one all-gather over it, so every rank returns
    the full (nq, k) answer like the reference's single host process does."""

    def __init__(self, inner=None, replicas: int | None = None, replica_id: int | None = None, cross_group=None):
        import torch.distributed as dist
        self.dist = dist
        self.inner = inner
        self.group = cross_group
        live = dist.is_initialized()
        if inner is None:       # the reference's manual form: IndexReplicas() then addIndex(...) (single process)
            replicas, replica_id = 1, 0
        self.replicas = replicas if replicas is not None else (dist.get_world_size(cross_group) if live else 1)
        self.replica_id = replica_id if replica_id is not None else (dist.get_rank(cross_group) if live else 0)
        self.d = getattr(inner, "d", None)
        self.peer_merge = getattr(inner, "peer_merge", False)
        self.peer_merge_error = getattr(inner, "peer_merge_error", None)
        self.own_fields = False
        self._members = [] if inner is None else [inner]

    def addIndex(self, index):
        """index = faiss.IndexReplicas(); index.addIndex(index1); index.own_fields = True
        (bench_gpu_performance_OSDI.py:613-626).  In one process every member is a copy of the same index on the same
        GPU, so the first one answers; across processes use the constructor arguments (one replica per rank)."""
        if not self._members:
            self.inner, self.d = index, getattr(index, "d", None)
        self._members.append(index)

    def count(self) -> int:
        return len(self._members)

    def at(self, i: int):
        return self._members[i]

    @property
    def ntotal(self):
        return getattr(self.inner, "ntotal", 0)

    @property
    def nprobe(self):
        return self.inner.nprobe

    @nprobe.setter
    def nprobe(self, v):
        self.inner.nprobe = v

    def search(self, xq: torch.Tensor, k: int):
        nq, R = xq.shape[0], self.replicas
        bounds = [(nq * r) // R for r in range(R + 1)]
        counts = [bounds[r + 1] - bounds[r] for r in range(R)]
        mine = xq[bounds[self.replica_id]:bounds[self.replica_id + 1]]
        if mine.shape[0]:
            D, I = self.inner.search(mine, k)
        else:   # fewer queries than replicas (the batch-1 latency path): nothing to do here but join the exchange
            D = torch.empty((0, k), dtype=torch.float32, device=xq.device)
            I = torch.empty((0, k), dtype=torch.int64, device=xq.device)
        if R == 1:
            return D, I
        return unpack_results(all_gather_rows(self.dist, self.group, R, pack_results(D, I), counts))


class DistributedIndexIVFPQ:
    """The rank-local view of a sharded index.

    search() = [coarse quantizer on this rank's slice of the batch -> all-gather of the probe lists] ->
               local scan of the whole batch against this rank's shard -> all-gather of (D, I) -> K5 merge.
    The codebooks are replicated, so every rank would compute the same probe lists; slicing the coarse stage by
    query removes that redundancy (it is the one stage whose cost does not shrink with the shard).  The probe
    exchange is nq * nprobe * 8 bytes per batch over NVLink.

    `merge_fn` / `local_search_fn` are injection points for the CPU (gloo) tests of the exchange logic; the
    product path uses the CUDA kernels."""

    def __init__(self, local_index, group=None, merge_fn=None, local_search_fn=None, shard_coarse=True, peer_merge=None,
                 shard_mode: str = "vector"):
        import torch.distributed as dist
        self.dist = dist
        self.local = local_index
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._merge = merge_fn or merge_shards
        self._injected_search = local_search_fn
        self.shard_coarse = shard_coarse
        self.d = getattr(local_index, "d", None)
        if shard_mode not in ("vector", "list", "replica"):
            raise ValueError("shard_mode must be 'vector', 'list' or 'replica'")
        # "list": this rank owns the lists l with l % world == rank (shard_index_by_list); probes of other ranks'
        # lists are masked out before the local scan.
        # "replica": every rank holds the WHOLE index (Faiss IndexReplicas, GpuMultipleClonerOptions.shard = False);
        # the batch is sliced by query, each rank answers its slice, one all-gather returns the full result everywhere
        self.shard_mode = shard_mode
        self._replicas = None
        if shard_mode == "replica":
            peer_merge = False
            inner = local_index if local_search_fn is None else _FnIndex(local_search_fn)
            self._replicas = IndexReplicas(inner, self.world, self.rank, group)
        if peer_merge is None and group is not None and dist.is_initialized() and group is not dist.group.WORLD:
            peer_merge = False      # symmetric memory over a subgroup (R x S layouts) is opt-in: not yet run on hardware
        # Peer-memory merge: every rank's (D, I) lands in a symmetric-memory buffer that all GPUs of the box map into
        # their address space; K5 then reads the shards in place over NVLink / NVSwitch -- no all-gather, no staging
        # copies.  Falls back to the NCCL all-gather when symmetric memory cannot be set up (or B200_IVFPQ_P2P=0).
        if peer_merge is None:
            peer_merge = os.environ.get("B200_IVFPQ_P2P", "1") != "0"
        self.peer_merge = bool(peer_merge) and merge_fn is None and local_search_fn is None and self.world > 1
        self.exchange_thresholds = (os.environ.get("B200_IVFPQ_THR_EXCHANGE", "1") != "0" and local_search_fn is None
                                    and merge_fn is None)
        self._symm = None          # (buffer, handle, capacity in result slots)
        self._readers_pending = False   # peers may still be reading my buffer (previous search's merge)
        self.peer_merge_error = None

    @property
    def nprobe(self):
        return self.local.nprobe

    @nprobe.setter
    def nprobe(self, v):
        self.local.nprobe = v

    def _all_gather_rows(self, mine: torch.Tensor, counts):
        return all_gather_rows(self.dist, self.group, self.world, mine, counts)

    def _local_search(self, xq: torch.Tensor, k: int, out=None):
        if self._injected_search is not None:
            return self._injected_search(xq, k)
        nq, nprobe = xq.shape[0], int(self.local.nprobe)
        sliced = self.shard_coarse and self.world > 1 and nq >= 8 * self.world
        if not sliced and self.shard_mode == "vector":
            return self.local.search(xq, k, out=out) if out is not None else self.local.search(xq, k)
        if sliced and self.shard_mode == "vector" and self.exchange_thresholds and hasattr(self.local, "prepare_queries") \
                and isinstance(xq, torch.Tensor) and xq.is_cuda and xq.dtype == torch.float32:
            # the filter tables depend on the queries only: they are built on a side stream while this rank's coarse
            # slice and the probe exchange run
            xq = self.local.prepare_queries(xq)
        if sliced:
            # coarse stage on my slice of the queries, then exchange the probe lists
            bounds = [(nq * r) // self.world for r in range(self.world + 1)]
            counts = [bounds[r + 1] - bounds[r] for r in range(self.world)]
            _, ids = self.local.quantizer.search(xq[bounds[self.rank]:bounds[self.rank + 1]],
                                                 min(nprobe, self.local.nlist))
            probes = self._all_gather_rows(ids, counts)
        else:
            _, probes = self.local.quantizer.search(xq, min(nprobe, self.local.nlist))
        if self.shard_mode == "list":
            probes = torch.where(probes % self.world == self.rank, probes, torch.full_like(probes, -1))
        if sliced and self.shard_mode == "vector" and self.exchange_thresholds and \
                hasattr(self.local, "search_preassigned_begin"):
            # every rank bootstraps the filter thresholds of ITS slice of the queries on its own shard (any shard's k-th
            # best distance bounds the global one); one all-reduce MIN of nq x 4 bytes gives every rank all of them
            thr = self.local.search_preassigned_begin(xq, k, probes, bounds[self.rank], bounds[self.rank + 1])
            if thr is not None:
                self.dist.all_reduce(thr, op=self.dist.ReduceOp.MIN, group=self.group)
                return self.local.search_preassigned_finish(thr, nq, k, out=out)
        if out is not None:
            return self.local.search_preassigned(xq, k, probes, out=out)
        return self.local.search_preassigned(xq, k, probes)

    # ---- peer-memory path --------------------------------------------------------------------------------
    def _symm_setup(self, slots: int):
        """(Re)allocate the symmetric result buffer for `slots` (= nq * k) results.  Collective: every rank calls it
        with the same size because every rank searches the same batch."""
        import torch.distributed._symmetric_memory as symm
        dev = torch.device("cuda", torch.cuda.current_device())
        cap = max(1 << 16, 1 << (int(slots) - 1).bit_length())
        group = self.group if self.group is not None else self.dist.group.WORLD
        if self._symm is not None and self._readers_pending:
            self._symm[1].barrier(channel=1)          # peers are done with the old buffer before it is dropped
            self._readers_pending = False
        buf = symm.empty((cap * 12 + 256,), dtype=torch.uint8, device=dev)
        hdl = symm.rendezvous(buf, group)
        self._symm = (buf, hdl, cap)

    def _search_peer(self, xq: torch.Tensor, k: int):
        nq = xq.shape[0]
        slots = nq * k
        if self._symm is None or self._symm[2] < slots:
            self._symm_setup(slots)
        buf, hdl, cap = self._symm
        d_off, i_off = 0, ((cap * 4 + 255) // 256) * 256
        D_loc = buf[d_off:d_off + slots * 4].view(torch.float32).view(nq, k)
        I_loc = buf[i_off:i_off + slots * 8].view(torch.int64).view(nq, k)
        if self._readers_pending:
            # nobody overwrites its buffer while a peer still reads it.  Enqueued here, i.e. in front of this search's
            # own milliseconds of scanning, the wait is over long before it matters
            hdl.barrier(channel=1)
            self._readers_pending = False
        self._local_search(xq, k, out=(D_loc, I_loc))
        hdl.barrier(channel=0)                       # every shard's results are in place (device-side, no host sync)
        D = torch.empty((nq, k), dtype=torch.float32, device=xq.device)
        I = torch.empty((nq, k), dtype=torch.int64, device=xq.device)
        lib = _lib.load()
        with torch.cuda.device(xq.device):
            st = int(torch.cuda.current_stream(xq.device).cuda_stream)
            _lib.check(lib.b200_ivfpq_merge_shards_peer(self.world, nq, k, int(hdl.buffer_ptrs_dev), d_off, i_off,
                                                        D.data_ptr(), I.data_ptr(), st))
        self._readers_pending = True
        return D, I

    def search(self, xq: torch.Tensor, k: int):
        if self._replicas is not None:
            return self._replicas.search(xq, k)
        if self.peer_merge and isinstance(xq, torch.Tensor) and xq.is_cuda:
            try:
                return self._search_peer(xq, k)
            except Exception as e:   # symmetric memory unavailable on this system: use the NCCL path from now on
                if self._symm is not None:
                    raise
                self.peer_merge = False
                self.peer_merge_error = f"{type(e).__name__}: {e}"
        D, I = self._local_search(xq, k)
        if self.world == 1:
            return D, I
        mine = pack_results(D, I)
        gathered = torch.empty((self.world * mine.shape[0],) + mine.shape[1:], dtype=mine.dtype, device=mine.device)
        self.dist.all_gather_into_tensor(gathered, mine, group=self.group)
        Ds, Is = unpack_results(gathered.view((self.world,) + mine.shape))
        return self._merge(Ds, Is)
