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
