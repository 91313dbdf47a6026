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


class AsyncB200Retriever(LocalB200Retriever):
    """In-process replacement for ExternalRetriever's split send / recv protocol
    (llm_inference_gpu/ralm/retriever/retriever.py:109-163), which the tik-tok decoder uses to overlap the retrieval
    of micro-batch A with the decode of micro-batch B (ralm/ralm_tiktok.py:129-192, 226-233).

    retrieve_send(query, k)  enqueues the search on a side CUDA stream and returns at once.  `query` may be the
                             decoder's hidden state as a CUDA tensor: no D2H copy (ralm.py:110-111 copies to the
                             CPU every retrieval step) and no TCP hop; the side stream first waits for the
                             producer stream, so the hidden state is complete before it is read.
    poll()                   non-blocking: True when the answer is ready (ExternalRetriever.poll polls a socket).
    retrieve_recv(k)         returns (indices, distances) like ExternalRetriever.retrieve_recv.  CUDA tensors in ->
                             CUDA tensors out, and the CALLER's stream waits on the event (no host sync);
                             numpy in -> numpy out.

    The index handle owns ONE workspace: while a send is pending, do not search the same index object from another
    stream (the two searches would share the probe table and candidate buffers).  Give the retriever its own index
    object (same codebook / list tensors, a second handle) if the decoder also searches.
    """

    def __init__(self, index: IndexIVFPQ, default_k: Optional[int] = 10, nprobe: Optional[int] = 1,
                 device: Optional[str] = "gpu"):
        self._dev = index._device()
        self._stream = torch.cuda.Stream(device=self._dev)
        self._event = None
        self._pending = None
        self._as_numpy = False
        super().__init__(index, default_k=default_k, nprobe=nprobe, device=device)   # runs the warm-up search

    def retrieve_send(self, query, k: Optional[int] = None, nprobe: Optional[int] = None):
        if self._pending is not None:
            raise RuntimeError("retrieve_send called twice without retrieve_recv")
        if k is None:
            k = self.default_k
        if nprobe is not None:
            self.set_nprobe(nprobe)
        self._as_numpy = not isinstance(query, torch.Tensor)
        producer = torch.cuda.current_stream(self._dev)
        ready = torch.cuda.Event()
        ready.record(producer)
        with torch.cuda.stream(self._stream):
            self._stream.wait_event(ready)
            if self._as_numpy:
                q = torch.from_numpy(np.ascontiguousarray(query, np.float32)).to(self._dev, non_blocking=True)
            else:
                q = query.to(self._dev).float().contiguous()
                q.record_stream(self._stream)
            D, I = self.index.search(q, k)
            self._event = torch.cuda.Event()
            self._event.record(self._stream)
        self._pending = (D, I, q)

    def poll(self) -> bool:
        return self._event is not None and self._event.query()

    def retrieve_recv(self, k: Optional[int] = None):
        if self._pending is None:
            raise RuntimeError("retrieve_recv without a pending retrieve_send")
        D, I, _ = self._pending
        self._pending = None
        if self._as_numpy:
            self._event.synchronize()
            return I.cpu().numpy(), D.cpu().numpy()
        consumer = torch.cuda.current_stream(self._dev)
        consumer.wait_event(self._event)
        # D and I were allocated under the side stream: tell the caching allocator that the caller's stream uses them,
        # or their blocks could be handed to the next side-stream search while the caller still reads them
        D.record_stream(consumer)
        I.record_stream(consumer)
        return I, D

    def retrieve(self, query, nprobe: Optional[int] = None, k: Optional[int] = None):
        self.retrieve_send(query, k=k, nprobe=nprobe)
        I, D = self.retrieve_recv(k)
        return {"id": I, "dist": D}
