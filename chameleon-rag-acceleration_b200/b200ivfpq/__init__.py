"""b200ivfpq -- B200-native IVF-PQ search engine with the Faiss-style Python surface Chameleon drives.

    import b200ivfpq as faiss
    index = faiss.index_factory(128, "IVF1024,PQ16")
    index.train(xt); index.add(xb); index.nprobe = 16
    D, I = index.search(xq, 10)

Hot path = hand-written sm_100a CUDA kernels behind a C-ABI (include/b200_ivfpq.h); no CPU fallback.
"""
from ._lib import LIB_PATH, launch_count, load as load_library
from . import contrib
from .compat import (INDICES_32_BIT, INDICES_64_BIT, INDICES_CPU, INDICES_IVF, Clustering, GpuClonerOptions,
                     GpuMultipleClonerOptions, GpuResourcesVector, IndexIVFFlat, IntVector, PCAMatrix, StandardGpuResources,
                     cvar, deserialize_index, float_maxheap_array_t,
                     get_num_gpus, index_cpu_to_all_gpus, index_cpu_to_gpu, index_cpu_to_gpu_multiple,
                     index_cpu_to_gpus_list, index_gpu_to_cpu, ranklist_intersection_size, rev_swig_ptr, serialize_index,
                     swig_ptr, vector_float_to_array)
from .factory import GpuParameterSpace, ParameterSpace, index_factory
from .index import METRIC_L2, IndexFlatL2, IndexIVFPQ, InvertedLists, ProductQuantizer
from .io import IO_FLAG_MMAP, IO_FLAG_ONDISK_SAME_DIR, IO_FLAG_READ_ONLY, read_index, write_index
from .retriever import AsyncB200Retriever, IndexScanner, LocalB200Retriever
from .server import B200Client, B200Server
from .transforms import (IndexPreTransform, OPQMatrix, VectorTransform, downcast_VectorTransform, read_VectorTransform,
                         write_VectorTransform)
from .shards import (DistributedIndexIVFPQ, IndexReplicas, make_replica_groups, merge_shards, replica_layout, shard_index,
                     shard_index_by_list, shard_positions)


def search_preassigned(index, xq, k, list_ids, coarse_dis=None):
    """faiss.contrib.ivf_tools.search_preassigned (ralm/server/faiss_server.py:233)."""
    return index.search_preassigned(xq, k, list_ids)


def omp_set_num_threads(n):   # accepted for source compatibility (faiss_retriever.py:79); nothing to set
    return None


def vector_to_array(v):
    import numpy as np
    return np.asarray(v)


def downcast_index(index):
    return index


__all__ = ["IndexFlatL2", "IndexIVFPQ", "index_factory", "ParameterSpace", "GpuParameterSpace", "search_preassigned",
           "read_index", "write_index", "LocalB200Retriever", "AsyncB200Retriever", "IndexScanner", "DistributedIndexIVFPQ", "IndexReplicas", "make_replica_groups", "replica_layout", "shard_index", "shard_index_by_list",
           "merge_shards", "B200Server", "B200Client", "IndexPreTransform", "OPQMatrix", "downcast_VectorTransform", "METRIC_L2", "omp_set_num_threads", "vector_to_array", "downcast_index", "launch_count",
           "StandardGpuResources", "GpuResourcesVector", "IntVector", "GpuClonerOptions", "GpuMultipleClonerOptions",
           "index_cpu_to_gpu", "index_cpu_to_gpu_multiple", "index_cpu_to_gpus_list", "index_cpu_to_all_gpus",
           "index_gpu_to_cpu", "serialize_index", "deserialize_index", "swig_ptr", "rev_swig_ptr", "cvar"]
