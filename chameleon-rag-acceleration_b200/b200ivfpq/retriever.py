"""Retriever adapter for the RAG serving loop: the drop-in for
llm_inference_gpu/ralm/retriever/faiss_retriever.py:18-275 (LocalFaissRetriever) and
llm_inference_gpu/ralm/index_scanner/index_scanner.py:15-77 (IndexScanner).

retrieve(query, nprobe, k) -> {"id": int64[nq, k], "dist": float32[nq, k]}  (faiss_retriever.py:227-275)
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .index import IndexFlatL2, IndexIVFPQ


class LocalB200Retriever:
    """Same constructor intent and retrieve() contract as LocalFaissRetriever, over an in-memory index.
    `device` strings of the reference ('cpu', 'gpu', 'cpu-gpu') are accepted; everything runs on the B200."""

    def __init__(self, index: IndexIVFPQ, default_k: Optional[int] = 10, nprobe: Optional[int] = 1,
                 device: Optional[str] = "gpu"):
        self.index = index
        self.dim = index.d
        self.default_k = default_k
        self.device = device
        self.set_nprobe(nprobe)
        # warm up search (faiss_retriever.py:81-83)
        if index.ntotal > 0:
            self.retrieve(np.random.rand(1, self.dim).astype("float32"), nprobe=1, k=1)
            self.set_nprobe(nprobe)

    def set_nprobe(self, nprobe: int):
        self.nprobe = nprobe
        self.index.nprobe = nprobe

    def retrieve(self, query, nprobe: Optional[int] = None, k: Optional[int] = None):
        if k is None:
            k = self.default_k
        if nprobe is None:
            nprobe = self.nprobe
        else:
            self.set_nprobe(nprobe)
        nq, dim = query.shape
        assert dim == self.dim
        D, I = self.index.search(query, k)
        return {"id": I, "dist": D}

    def retrieve_with_lists(self, query, list_IDs, k: Optional[int] = None):
        """faiss_server.py:220-239."""
        if k is None:
            k = self.default_k
        nq, dim = query.shape
        assert dim == self.dim
        assert list_IDs.shape[0] == nq
        D, I = self.index.search_preassigned(query, k, list_IDs)
        return {"id": I, "dist": D}


class IndexScanner:
    """index_scanner.py:15-77: coarse scan only; returns list ids and their centroid vectors."""

    def __init__(self, dim: int = 1024, nlist: int = 32768, nprobe: int = 32, centroids=None, device: str = "gpu",
                 use_gpu_id: Optional[int] = None, omp_threads: Optional[int] = None):
        self.dim, self.nlist, self.nprobe = dim, nlist, nprobe
        if centroids is None:
            centroids = np.random.rand(nlist, dim).astype("float32")
        assert centroids.shape == (nlist, dim)
        self.centroids = centroids
        if use_gpu_id is not None:
            torch.cuda.set_device(use_gpu_id)
        self.index = IndexFlatL2(dim)
        self.index.add(centroids)
        self.index.search(np.random.rand(1, dim).astype("float32"), nprobe)

    def search(self, queries, nprobe: Optional[int] = None):
        assert queries.shape[1] == self.dim
        nq = queries.shape[0]
        if nprobe is None:
            nprobe = self.nprobe
        D, I = self.index.search(queries, nprobe)
        list_IDs = np.array(I, dtype="int64")
        list_centroids = self.centroids[list_IDs.flatten()].reshape(nq, nprobe, self.dim)
        return list_IDs, list_centroids
