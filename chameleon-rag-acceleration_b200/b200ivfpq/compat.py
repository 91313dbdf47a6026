"""Faiss names the reference's drivers call around the hot path, mapped onto this engine.

These are thin: an index built by this package already lives on the B200, so the "CPU -> GPU" cloners return the
index itself -- or, inside a torchrun job with `co.shard = True`, this rank's shard wrapped in DistributedIndexIVFPQ,
which is exactly what faiss.index_cpu_to_gpu_multiple(vres, vdev, index, co) builds with one process per GPU
(bench_gpu_performance_OSDI.py:586-604; faiss_retriever.py:120 index_cpu_to_gpus_list).
"""
from __future__ import annotations

import io as _io

import numpy as np


# ---- SWIG pointer helpers: get_ids / get_codes / get_xb already return numpy arrays here ------------------------------
def swig_ptr(a):
    """faiss.swig_ptr(array): arrays are passed as they are."""
    return a


def rev_swig_ptr(ptr, n: int):
    """faiss.rev_swig_ptr(ptr, n) (extract_Enzian_U250_required_data.py:264-279): first n elements as a numpy view."""
    return np.asarray(ptr).reshape(-1)[:int(n)]


def get_num_gpus() -> int:
    """faiss.get_num_gpus() (bench_gpu_performance_OSDI.py:109, faiss_retriever.py:153)."""
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def vector_float_to_array(v):
    """faiss.vector_float_to_array(clus.centroids) (bench_gpu_1bn.py:540)."""
    return np.asarray(v, dtype=np.float32).reshape(-1)


def ranklist_intersection_size(k1: int, v1, k2: int, v2) -> int:
    """faiss.ranklist_intersection_size(k1, swig_ptr(a), k2, swig_ptr(b)) (bench_gpu_1bn.py:222): how many of the first
    k1 ids of a are among the first k2 ids of b."""
    a = np.asarray(v1).reshape(-1)[:int(k1)]
    b = np.asarray(v2).reshape(-1)[:int(k2)]
    return int(np.intersect1d(a, b).shape[0])


# GpuClonerOptions.indicesOptions values (bench_gpu_1bn.py:608); ids always live next to the codes in HBM here
INDICES_CPU, INDICES_IVF, INDICES_32_BIT, INDICES_64_BIT = 0, 1, 2, 3


class Clustering:
    """faiss.Clustering(d, k) as the reference trains its coarse quantizer with it (bench_gpu_1bn.py:520-542):
        clus = faiss.Clustering(d, k); clus.verbose = True; clus.max_points_per_centroid = 10000000
        clus.train(x, index); centroids = faiss.vector_float_to_array(clus.centroids).reshape(k, d)
    Lloyd iterations on the GPU (b200ivfpq.kmeans); there is no CPU path: without a CUDA device this raises like the
    rest of the package.  `index` receives the centroids like Faiss's assignment index does."""

    def __init__(self, d: int, k: int):
        self.d, self.k = int(d), int(k)
        self.niter, self.seed = 25, 1234
        self.max_points_per_centroid, self.min_points_per_centroid = 256, 39
        self.verbose = False
        self.centroids = np.zeros(0, np.float32)

    def train(self, x, index=None):
        import torch
        from .kmeans import kmeans
        x = np.ascontiguousarray(x, np.float32) if not isinstance(x, torch.Tensor) else x
        xt = torch.as_tensor(x, dtype=torch.float32)
        if xt.dim() != 2 or xt.shape[1] != self.d:
            raise AssertionError(f"Clustering.train: expected (n, {self.d}) vectors")
        if not torch.cuda.is_available():
            raise RuntimeError("b200ivfpq needs a CUDA device (B200, sm_100a); there is no CPU path")
        xt = xt.cuda()
        c = kmeans(xt, self.k, niter=self.niter, seed=self.seed, max_points_per_centroid=self.max_points_per_centroid,
                   verbose=self.verbose)
        self.centroids = c.reshape(-1).cpu().numpy()
        if index is not None:
            index.reset()
            index.add(c)


class float_maxheap_array_t:
    """faiss.float_maxheap_array_t as the reference's ground-truth loop uses it (bench_gpu_1bn.py:427-456): nh rows of
    the k smallest (value, id) seen so far, living IN the caller's arrays:
        heaps.k, heaps.nh = k, nq; heaps.val = swig_ptr(D); heaps.ids = swig_ptr(I); heaps.heapify()
        heaps.addn_with_ids(k, swig_ptr(D_block), swig_ptr(I_block), k)   # per block of the database
        heaps.reorder()                                                   # rows ascending
    A value replaces the current worst only if it is strictly smaller (entries already held win ties)."""

    def __init__(self):
        self.k = self.nh = 0
        self.val = self.ids = None

    def _views(self):
        val = np.asarray(self.val).reshape(self.nh, self.k)
        ids = np.asarray(self.ids).reshape(self.nh, self.k)
        if not (np.shares_memory(val, self.val) and np.shares_memory(ids, self.ids)):
            raise RuntimeError("float_maxheap_array_t: val / ids must be C-contiguous arrays (they are updated in place)")
        return val, ids

    def heapify(self):
        val, ids = self._views()
        val[:] = np.finfo(np.float32).max
        ids[:] = -1

    def addn_with_ids(self, nj: int, vin, id_in, id_stride: int = 0, i0: int = 0, ni: int = -1):
        val, ids = self._views()
        ni = self.nh - i0 if ni < 0 else ni
        vin = np.asarray(vin, np.float32).reshape(ni, nj)
        id_in = np.asarray(id_in, np.int64).reshape(ni, id_stride or nj)[:, :nj]
        allv = np.concatenate([val[i0:i0 + ni], vin], axis=1)
        alli = np.concatenate([ids[i0:i0 + ni], id_in], axis=1)
        order = np.argsort(allv, axis=1, kind="stable")[:, :self.k]       # held entries come first among equals
        val[i0:i0 + ni] = np.take_along_axis(allv, order, axis=1)
        ids[i0:i0 + ni] = np.take_along_axis(alli, order, axis=1)

    def reorder(self):
        val, ids = self._views()
        order = np.argsort(val, axis=1, kind="stable")
        val[:] = np.take_along_axis(val, order, axis=1)
        ids[:] = np.take_along_axis(ids, order, axis=1)


class PCAMatrix:
    """Named for completeness of the drivers' imports (bench_gpu_1bn.py `PCAR<d>` keys); not built: the reference's
    hot-path configurations use OPQ or no pre-transform (BASELINE.json)."""

    def __init__(self, *args, **kwargs):
        raise RuntimeError("PCAMatrix is not supported; use OPQ<m>[_<d>] or no pre-transform")


class IndexIVFFlat:
    """Present so that `index.__class__ == faiss.IndexIVFFlat` tests (bench_gpu_1bn.py:692) evaluate; building one is out
    of scope (this engine is IVF-PQ)."""

    def __init__(self, *args, **kwargs):
        raise RuntimeError("IndexIVFFlat is not supported: this engine implements IVF-PQ (IVF<nlist>,PQ<m>)")


# ---- GPU resources / cloner options: accepted, nothing to configure ------------------------------------------------
class StandardGpuResources:
    def setTempMemory(self, nbytes):
        self.temp_memory = int(nbytes)

    def setPinnedMemory(self, nbytes):
        self.pinned_memory = int(nbytes)

    def noTempMemory(self):
        self.temp_memory = 0

    def setDefaultNullStreamAllDevices(self):
        pass

    def syncDefaultStreamCurrentDevice(self):
        import torch
        torch.cuda.current_stream().synchronize()


class _Vector(list):
    """faiss.GpuResourcesVector / faiss.IntVector: push_back + len."""

    def push_back(self, v):
        self.append(v)

    def size(self):
        return len(self)

    def at(self, i):
        return self[i]


GpuResourcesVector = _Vector
IntVector = _Vector


class GpuClonerOptions:
    def __init__(self):
        self.useFloat16 = False
        self.useFloat16CoarseQuantizer = False
        self.usePrecomputed = False
        self.indicesOptions = 0
        self.reserveVecs = 0
        self.verbose = False


class GpuMultipleClonerOptions(GpuClonerOptions):
    def __init__(self):
        super().__init__()
        self.shard = False
        self.shard_type = 1


def _maybe_shard(index, co):
    """Single process: the index is already on the GPU.  torchrun job + co.shard: this rank's modulo shard behind the
    all-gather / peer-memory merge; torchrun job without co.shard: a replica that answers its slice of every batch."""
    import torch.distributed as dist
    from .shards import DistributedIndexIVFPQ, shard_index
    if co is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        inner = getattr(index, "index", index)
        if getattr(co, "shard", False):
            return DistributedIndexIVFPQ(shard_index(inner, dist.get_rank(), dist.get_world_size()))
        # co.shard = False is Faiss's IndexReplicas: full copy per GPU, queries split between them
        return DistributedIndexIVFPQ(inner, shard_mode="replica")
    return index


def index_cpu_to_gpu(res, device, index, options=None):
    return index


def index_cpu_to_gpu_multiple(vres, vdev, index, co=None):
    return _maybe_shard(index, co)


def index_cpu_to_gpus_list(index, co=None, gpus=None, ngpu=-1):
    return _maybe_shard(index, co)


def index_cpu_to_all_gpus(index, co=None, ngpu=-1):
    return _maybe_shard(index, co)


def index_gpu_to_cpu(index):
    return index


# ---- serialisation to bytes (faiss.serialize_index returns a uint8 numpy array) ---------------------------------------
def serialize_index(index) -> np.ndarray:
    from .io import write_index
    import os
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        fn = os.path.join(tmp, "index.npz")
        write_index(index, fn)
        return np.frombuffer(open(fn, "rb").read(), dtype=np.uint8).copy()


def deserialize_index(data):
    from .io import read_index
    import os
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        fn = os.path.join(tmp, "index.npz")
        with open(fn, "wb") as f:
            f.write(np.asarray(data, np.uint8).tobytes())
        return read_index(fn)


# ---- statistics objects the drivers reset and print (bench_cpu_performance.py, profiling_stages.py) -----------------
class _Stats:
    _fields = ("nq", "nlist", "ndis", "nheap_updates", "quantization_time", "search_time", "n_hamming_pass",
               "search_cycles", "refine_cycles")

    def __init__(self):
        self.reset()

    def reset(self):
        for f in self._fields:
            setattr(self, f, 0)


class _CVar:
    def __init__(self):
        self.indexIVF_stats = _Stats()
        self.indexIVFPQ_stats = _Stats()


cvar = _CVar()
