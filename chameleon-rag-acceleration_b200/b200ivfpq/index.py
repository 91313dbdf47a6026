"""Faiss-style index objects backed by the B200 CUDA library.

The surface is the one Chameleon's drivers use on Faiss:
  Faiss_experiments/IVFPQ_random_dataset.py:20-46   IndexFlatL2(d); IndexIVFPQ(quantizer, d, nlist, m, nbits);
                                                    index.train / add / search(xq, k); index.nprobe
  Faiss_experiments/bench_cpu_performance.py:98-269 index_factory, add in blocks, ParameterSpace, search loop
  my_faiss_extract_scripts/extract_Enzian_U250_required_data.py:222-279
                                                    pq.M/ksub/dsub/centroids, quantizer, invlists.list_size /
                                                    get_ids / get_codes / code_size
  llm_inference_gpu/ralm/retriever/faiss_retriever.py:227-275  retrieve -> index.search / search_preassigned

numpy in -> numpy out goes through the C-ABI's host entry point (H2D + search + D2H).  torch CUDA tensors
in -> torch CUDA tensors out stays on the device and on the caller's current stream.
All data lives in HBM as torch tensors; the C library borrows their pointers.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np
import torch

from . import _lib
from .kmeans import kmeans, kmeans_subspaces

METRIC_L2 = 1
MAX_K = 2048
MAX_NPROBE = 2048


def _require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("b200ivfpq needs a CUDA device (B200, sm_100a); there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def _as_f32_matrix(x, d: int, what: str):
    """Faiss asserts on shape and dtype (the scripts call sanitize/astype('float32') first)."""
    if isinstance(x, torch.Tensor):
        if x.dim() != 2 or x.shape[1] != d:
            raise AssertionError(f"{what}: expected shape (n, {d}), got {tuple(x.shape)}")
        if x.dtype != torch.float32:
            raise TypeError(f"{what}: expected float32, got {x.dtype}")
        return x.contiguous()
    x = np.asarray(x)
    if x.ndim != 2 or x.shape[1] != d:
        raise AssertionError(f"{what}: expected shape (n, {d}), got {x.shape}")
    if x.dtype != np.float32:
        raise TypeError(f"{what}: expected float32, got {x.dtype}")
    return np.ascontiguousarray(x)


def _check_out(out, nq: int, k: int, dev):
    D, I = out
    ok = (isinstance(D, torch.Tensor) and isinstance(I, torch.Tensor) and D.is_cuda and I.is_cuda and
          D.dtype == torch.float32 and I.dtype == torch.int64 and tuple(D.shape) == (nq, k) and
          tuple(I.shape) == (nq, k) and D.is_contiguous() and I.is_contiguous() and D.device == dev and I.device == dev)
    if not ok:
        raise AssertionError("out = (D, I) must be contiguous CUDA tensors (nq, k) f32 / i64 on the index's device")
    return D, I


def _to_device(x, device) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x.to(device, non_blocking=True).contiguous()
    return torch.from_numpy(x).to(device)


def _stream_ptr(device) -> int:
    return int(torch.cuda.current_stream(device).cuda_stream)


class _Handle:
    """Owns one b200_ivfpq_t."""

    def __init__(self, d: int, nlist: int, m: int, nbits: int = 8):
        self.lib = _lib.load()
        h = ctypes.c_void_p()
        _lib.check(self.lib.b200_ivfpq_create(d, nlist, m, nbits, ctypes.byref(h)))
        self.h = h

    def __del__(self):
        try:
            if getattr(self, "h", None) and self.h.value:
                self.lib.b200_ivfpq_destroy(self.h)
                self.h = ctypes.c_void_p()
        except Exception:
            pass


class _SwigHandle:
    """index.this.disown() / .own(): SWIG ownership calls the drivers make before handing an index to a container
    (bench_gpu_performance_OSDI.py:624); Python owns everything here."""

    def disown(self):
        return None

    def own(self):
        return None


class IndexFlatL2:
    """Exact L2 index (the coarse quantizer, and brute-force ground truth).

    search() runs kernel K1 (exact fp32 distances + select) over the stored vectors.
    Reference: quantizer = faiss.IndexFlatL2(d) (IVFPQ_random_dataset.py:22); IndexScanner
    (llm_inference_gpu/ralm/index_scanner/index_scanner.py:33-73) adds centroids and searches nprobe.
    """
    this = _SwigHandle()

    def __init__(self, d: int):
        self.d = int(d)
        self.ntotal = 0
        self.is_trained = True
        self.metric_type = METRIC_L2
        self._xb: Optional[torch.Tensor] = None
        self._handle: Optional[_Handle] = None

    def train(self, x):
        return None

    def reset(self):
        self._xb, self._handle, self.ntotal = None, None, 0

    def add(self, x):
        x = _as_f32_matrix(x, self.d, "add")
        dev = _require_cuda()
        xt = _to_device(x, dev)
        self._xb = xt.clone() if self._xb is None else torch.cat([self._xb, xt], 0)
        self.ntotal = int(self._xb.shape[0])
        self._handle = None

    def _set_xb(self, xb: torch.Tensor):
        self._xb = xb.contiguous()
        self.ntotal = int(xb.shape[0])
        self._handle = None

    def xb_tensor(self) -> torch.Tensor:
        return self._xb

    def get_xb(self) -> np.ndarray:
        """faiss.rev_swig_ptr(quantizer.get_xb(), nlist * d) equivalent, already shaped (ntotal, d)."""
        return self._xb.cpu().numpy() if self._xb is not None else np.zeros((0, self.d), np.float32)

    def reconstruct_n(self, i0: int, n: int) -> np.ndarray:
        return self._xb[i0:i0 + n].cpu().numpy()

    def _ensure_handle(self):
        if self.ntotal == 0:
            raise RuntimeError("IndexFlatL2 is empty")
        if self._handle is None:
            h = _Handle(self.d, self.ntotal, 1, 8)
            _lib.check(h.lib.b200_ivfpq_set_codebooks(h.h, self._xb.data_ptr(), None))
            self._handle = h
        return self._handle

    def search(self, x, k: int):
        x = _as_f32_matrix(x, self.d, "search")
        if not (1 <= k <= MAX_K):
            raise RuntimeError(f"k = {k} out of [1, {MAX_K}]")
        h = self._ensure_handle()
        dev = self._xb.device
        is_torch = isinstance(x, torch.Tensor)
        xq = _to_device(x, dev)
        nq = xq.shape[0]
        D = torch.empty((nq, k), dtype=torch.float32, device=dev)
        I = torch.empty((nq, k), dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(h.lib.b200_ivfpq_coarse(h.h, nq, xq.data_ptr(), k, I.data_ptr(), D.data_ptr(), _stream_ptr(dev)))
        if is_torch:
            return D, I
        return D.cpu().numpy(), I.cpu().numpy()


class ProductQuantizer:
    """index.pq view: M, ksub, dsub, nbits, centroids (flat, as faiss.vector_to_array(pq.centroids))."""

    def __init__(self, d: int, M: int, nbits: int):
        self.d, self.M, self.nbits = d, M, nbits
        self.ksub = 1 << nbits
        self.dsub = d // M
        self.code_size = M * nbits // 8
        self._centroids: Optional[torch.Tensor] = None     # (M, ksub, dsub) on the GPU

    @property
    def centroids(self) -> np.ndarray:
        if self._centroids is None:
            return np.zeros(0, np.float32)
        return self._centroids.reshape(-1).cpu().numpy()

    def centroids_tensor(self) -> torch.Tensor:
        return self._centroids


class InvertedLists:
    """index.invlists view (Faiss ArrayInvertedLists): nlist, code_size, list_size(l), get_ids(l), get_codes(l).
    get_ids / get_codes return numpy copies (the reference wraps raw pointers with rev_swig_ptr and then copies,
    extract_Enzian_U250_required_data.py:264-279)."""

    def __init__(self, index: "IndexIVFPQ"):
        self._index = index

    @property
    def nlist(self) -> int:
        return self._index.nlist

    @property
    def code_size(self) -> int:
        return self._index.pq.code_size

    def list_size(self, l: int) -> int:
        off = self._index._finalized_offsets()
        return int(off[l + 1] - off[l])

    def get_ids(self, l: int) -> np.ndarray:
        off = self._index._finalized_offsets()
        if off[l + 1] == off[l]:
            return np.zeros(0, np.int64)
        return self._index._ids[off[l]:off[l + 1]].cpu().numpy()

    def get_codes(self, l: int) -> np.ndarray:
        """uint8 array of list_size * code_size bytes (flat, like rev_swig_ptr(get_codes(l), ls * code_size))."""
        off = self._index._finalized_offsets()
        if off[l + 1] == off[l]:
            return np.zeros(0, np.uint8)
        return self._index._codes[off[l]:off[l + 1]].reshape(-1).cpu().numpy()

    def imbalance_factor(self) -> float:
        off = self._index._finalized_offsets()
        sizes = np.diff(off).astype(np.float64)
        tot = sizes.sum()
        return float((sizes ** 2).sum() * len(sizes) / (tot * tot)) if tot else 0.0


class IndexIVFPQ:
    """IVF-PQ index: faiss.IndexIVFPQ(quantizer, d, nlist, m, nbits) (IVFPQ_random_dataset.py:24)."""
    this = _SwigHandle()

    def __init__(self, quantizer: IndexFlatL2, d: int, nlist: int, m: int, nbits: int = 8, metric=METRIC_L2):
        if nbits != 8:
            raise RuntimeError(f"nbits = {nbits}: only 8-bit PQ codes are supported (the reference only uses 8)")
        if d % m != 0:
            raise RuntimeError(f"The dimension of the vectors (d = {d}) should be a multiple of the number of "
                               f"subquantizers (M = {m})")
        if metric != METRIC_L2:
            raise RuntimeError("only METRIC_L2 is supported")
        self.d, self.nlist = int(d), int(nlist)
        self.quantizer = quantizer if quantizer is not None else IndexFlatL2(d)
        self.pq = ProductQuantizer(self.d, int(m), nbits)
        self.nprobe = 1
        self.ntotal = 0
        self.is_trained = False
        self.by_residual = True
        self.metric_type = METRIC_L2
        self.parallel_mode = 0          # accepted and ignored (faiss_retriever.py:71 sets 3)
        self.verbose = False
        self.cp_niter = 25              # ClusteringParameters.niter
        self.invlists = InvertedLists(self)
        # CSR inverted lists on the GPU
        self._codes: Optional[torch.Tensor] = None      # (ntotal, M) uint8, list-major
        self._ids: Optional[torch.Tensor] = None        # (ntotal,) int64
        self._offsets = np.zeros(self.nlist + 1, np.int64)
        self._pending = []                              # [(list_no i64, codes u8, ids i64)] not yet merged
        self._handle: Optional[_Handle] = None
        self._lists_dirty = True

    # ------------------------------------------------------------------ handle / state
    def _ensure_handle(self) -> _Handle:
        if self._handle is None:
            self._handle = _Handle(self.d, self.nlist, self.pq.M, self.pq.nbits)
            self._lists_dirty = True
            self._pushed_xb = None
            if self.is_trained:
                self._push_codebooks()
        elif self.is_trained and self.quantizer.xb_tensor() is not None and \
                getattr(self, "_pushed_xb", None) != self.quantizer.xb_tensor().data_ptr():
            # quantizer.add() / reset() replaced the centroid tensor: the library holds a borrowed pointer to the old one
            self._push_codebooks()
        return self._handle

    def _push_codebooks(self):
        h = self._handle
        if tuple(self.quantizer.xb_tensor().shape) != (self.nlist, self.d):
            raise RuntimeError(f"quantizer holds {tuple(self.quantizer.xb_tensor().shape)} vectors, expected "
                               f"({self.nlist}, {self.d})")
        _lib.check(h.lib.b200_ivfpq_set_codebooks(h.h, self.quantizer.xb_tensor().data_ptr(),
                                                  self.pq._centroids.data_ptr()))
        self._pushed_xb = self.quantizer.xb_tensor().data_ptr()

    def _device(self):
        if self.quantizer.xb_tensor() is not None:
            return self.quantizer.xb_tensor().device
        return _require_cuda()

    def set_codebooks(self, coarse_centroids, pq_centroids):
        """Install trained codebooks: coarse (nlist, d) and pq (M, 256, dsub) -- the arrays the reference
        extracts with get_coarse_quantizer_centroids / get_sub_quantizer_centroids
        (extract_Enzian_U250_required_data.py:222-246)."""
        dev = _require_cuda()
        c = _to_device(np.ascontiguousarray(coarse_centroids, np.float32)
                       if not isinstance(coarse_centroids, torch.Tensor) else coarse_centroids, dev)
        p = _to_device(np.ascontiguousarray(pq_centroids, np.float32)
                       if not isinstance(pq_centroids, torch.Tensor) else pq_centroids, dev)
        if tuple(c.shape) != (self.nlist, self.d):
            raise AssertionError(f"coarse centroids must be ({self.nlist}, {self.d}), got {tuple(c.shape)}")
        p = p.reshape(self.pq.M, self.pq.ksub, self.pq.dsub).contiguous()
        self.quantizer.reset()
        self.quantizer._set_xb(c.float())
        self.pq._centroids = p.float()
        self.is_trained = True
        self._ensure_handle()
        self._push_codebooks()

    def set_lists(self, offsets, codes, ids=None):
        """Install populated inverted lists in the flattened ArrayInvertedLists layout: list l = rows
        [offsets[l], offsets[l+1]) of codes (ntotal, M) uint8 and ids (ntotal,) int64."""
        dev = self._device()
        offsets = np.ascontiguousarray(offsets, np.int64)
        if offsets.shape != (self.nlist + 1,):
            raise AssertionError("offsets must have nlist + 1 entries")
        ntotal = int(offsets[-1])
        codes_t = codes if isinstance(codes, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(codes, np.uint8))
        codes_t = codes_t.to(dev).reshape(ntotal, self.pq.M).contiguous()
        if ids is None:
            ids_t = torch.arange(ntotal, dtype=torch.int64, device=dev)
        else:
            ids_t = ids if isinstance(ids, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(ids, np.int64))
            ids_t = ids_t.to(dev).contiguous()
        self._codes, self._ids, self._offsets = codes_t, ids_t, offsets.copy()
        self._pending = []
        self.ntotal = ntotal
        self._lists_dirty = True

    # ------------------------------------------------------------------ train
    def train(self, x):
        """k-means for the coarse quantizer (if it is empty), then per-subspace k-means on residuals."""
        x = _as_f32_matrix(x, self.d, "train")
        dev = _require_cuda()
        xt = _to_device(x, dev)
        if self.is_trained:
            return
        if self.quantizer.ntotal == 0:
            if xt.shape[0] < self.nlist:
                raise RuntimeError(f"Number of training points ({xt.shape[0]}) should be at least as large as "
                                   f"number of clusters ({self.nlist})")
            cent = kmeans(xt, self.nlist, niter=self.cp_niter, seed=1234, verbose=self.verbose)
            self.quantizer._set_xb(cent)
        elif self.quantizer.ntotal != self.nlist:
            raise RuntimeError("quantizer.ntotal != nlist")
        cent = self.quantizer.xb_tensor()
        # PQ training set: at most 256 points per sub-centroid, residuals w.r.t. the nearest centroid
        n_pq = min(xt.shape[0], self.pq.ksub * 256)
        g = torch.Generator(device=dev)
        g.manual_seed(999)
        sel = torch.randperm(xt.shape[0], generator=g, device=dev)[:n_pq]
        xs = xt[sel]
        from .kmeans import _assign
        labels, _ = _assign(xs, cent)
        res = xs - cent[labels]
        self.pq._centroids = kmeans_subspaces(res, self.pq.M, self.pq.ksub, niter=self.cp_niter, seed=4321)
        self.is_trained = True
        self._ensure_handle()
        self._push_codebooks()

    # ------------------------------------------------------------------ add
    def add(self, x):
        n = x.shape[0]
        ids = torch.arange(self.ntotal, self.ntotal + n, dtype=torch.int64, device=self._device())
        self.add_with_ids(x, ids)

    def add_with_ids(self, x, ids):
        """assign -> residual -> PQ encode on the GPU (b200_ivfpq_assign_encode), append to the lists."""
        if not self.is_trained:
            raise RuntimeError("Error: 'is_trained' failed (add called before train)")
        x = _as_f32_matrix(x, self.d, "add")
        dev = self._device()
        n = x.shape[0]
        if isinstance(ids, torch.Tensor):
            ids_t = ids.to(dev, torch.int64).contiguous()
        else:
            ids_t = torch.from_numpy(np.ascontiguousarray(ids, np.int64)).to(dev)
        if ids_t.shape != (n,):
            raise AssertionError("ids must have shape (n,)")
        h = self._ensure_handle()
        chunk = 1 << 21
        for i0 in range(0, n, chunk):
            xb = _to_device(x[i0:i0 + chunk], dev)
            nb = xb.shape[0]
            list_no = torch.empty(nb, dtype=torch.int64, device=dev)
            codes = torch.empty((nb, self.pq.M), dtype=torch.uint8, device=dev)
            if not bool(torch.isfinite(xb).all()):
                # the coarse quantizer answers id -1 for rows with NaN / inf distances: nothing to encode against
                raise RuntimeError("add: the vectors contain NaN or inf")
            with torch.cuda.device(dev):
                _lib.check(h.lib.b200_ivfpq_assign_encode(h.h, nb, xb.data_ptr(), list_no.data_ptr(), codes.data_ptr(),
                                                          _stream_ptr(dev)))
            self._pending.append((list_no.to(torch.int32), codes, ids_t[i0:i0 + nb]))
        self.ntotal += n
        self._lists_dirty = True

    def _finalize_lists(self):
        """Merge pending adds into the CSR arrays.  Order inside a list = insertion order (stable sort)."""
        if not self._pending:
            return
        dev = self._device()
        parts_l, parts_c, parts_i = [], [], []
        if self._codes is not None and self._codes.shape[0] > 0:
            sizes = torch.from_numpy(np.diff(self._offsets)).to(dev)
            parts_l.append(torch.repeat_interleave(torch.arange(self.nlist, dtype=torch.int32, device=dev), sizes))
            parts_c.append(self._codes)
            parts_i.append(self._ids)
        for l, c, i in self._pending:
            parts_l.append(l)
            parts_c.append(c)
            parts_i.append(i)
        self._pending = []
        list_no = torch.cat(parts_l)
        _, order = torch.sort(list_no, stable=True)
        counts = torch.bincount(list_no.long(), minlength=self.nlist)
        del list_no
        codes = torch.cat(parts_c)
        del parts_c
        self._codes = codes[order].contiguous()
        del codes
        ids = torch.cat(parts_i)
        self._ids = ids[order].contiguous()
        del ids, order
        off = np.zeros(self.nlist + 1, np.int64)
        off[1:] = np.cumsum(counts.cpu().numpy())
        self._offsets = off
        self._lists_dirty = True

    def _finalized_offsets(self) -> np.ndarray:
        self._finalize_lists()
        return self._offsets

    def _sync_lists(self):
        self._finalize_lists()
        h = self._ensure_handle()
        if self._lists_dirty:
            if self._codes is None:
                dev = self._device()
                self._codes = torch.zeros((0, self.pq.M), dtype=torch.uint8, device=dev)
                self._ids = torch.zeros((0,), dtype=torch.int64, device=dev)
            _lib.check(h.lib.b200_ivfpq_set_lists(h.h, self._offsets.ctypes.data, self._codes.data_ptr() or None,
                                                  self._ids.data_ptr() or None, int(self._offsets[-1])))
            self._lists_dirty = False
        return h

    # ------------------------------------------------------------------ search
    def _check_search(self, k: int, nprobe: int):
        if not self.is_trained:
            raise RuntimeError("Error: 'is_trained' failed (search called before train)")
        if not (1 <= k <= MAX_K):
            raise RuntimeError(f"k = {k} out of [1, {MAX_K}]")
        if not (1 <= nprobe <= MAX_NPROBE):
            raise RuntimeError(f"nprobe = {nprobe} out of [1, {MAX_NPROBE}]")

    def search(self, x, k: int, out=None):
        """index.search(xq, k) -> (D, I), rows ascending by distance, unfilled slots I = -1 / D = FLT_MAX.
        out = (D, I): preallocated CUDA tensors to write into (torch input only; used by the multi-GPU layer to have
        the results land directly in peer-visible memory)."""
        x = _as_f32_matrix(x, self.d, "search")
        nprobe = int(self.nprobe)
        self._check_search(k, nprobe)
        h = self._sync_lists()
        dev = self._device()
        nq = x.shape[0]
        with torch.cuda.device(dev):
            if isinstance(x, torch.Tensor):
                xq = x.to(dev).contiguous()
                if out is not None:
                    D, I = _check_out(out, nq, k, dev)
                else:
                    D = torch.empty((nq, k), dtype=torch.float32, device=dev)
                    I = torch.empty((nq, k), dtype=torch.int64, device=dev)
                _lib.check(h.lib.b200_ivfpq_search(h.h, nq, xq.data_ptr(), k, nprobe, D.data_ptr(), I.data_ptr(),
                                                   _stream_ptr(dev)))
                return D, I
            D = np.empty((nq, k), np.float32)
            I = np.empty((nq, k), np.int64)
            _lib.check(h.lib.b200_ivfpq_search_host(h.h, nq, x.ctypes.data, k, nprobe, D.ctypes.data, I.ctypes.data))
            return D, I

    def search_preassigned(self, x, k: int, list_ids, out=None):
        """faiss.contrib.ivf_tools.search_preassigned(index, xq, k, list_ids) (faiss_server.py:233)."""
        x = _as_f32_matrix(x, self.d, "search_preassigned")
        is_torch = isinstance(x, torch.Tensor)
        dev = self._device()
        if isinstance(list_ids, torch.Tensor):
            lids = list_ids.to(dev, torch.int64).contiguous()
        else:
            lids = torch.from_numpy(np.ascontiguousarray(list_ids, np.int64)).to(dev)
        nq = x.shape[0]
        if lids.dim() != 2 or lids.shape[0] != nq:
            raise AssertionError("list_ids must have shape (nq, nprobe)")
        nprobe = int(lids.shape[1])
        self._check_search(k, nprobe)
        h = self._sync_lists()
        xq = _to_device(x, dev)
        if out is not None and is_torch:
            D, I = _check_out(out, nq, k, dev)
        else:
            D = torch.empty((nq, k), dtype=torch.float32, device=dev)
            I = torch.empty((nq, k), dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(h.lib.b200_ivfpq_search_preassigned(h.h, nq, xq.data_ptr(), k, nprobe, lids.data_ptr(),
                                                           D.data_ptr(), I.data_ptr(), _stream_ptr(dev)))
        if is_torch:
            return D, I
        return D.cpu().numpy(), I.cpu().numpy()

    def prepare_queries(self, x: torch.Tensor) -> torch.Tensor:
        """Optional head start for the next search of exactly these queries (b200_ivfpq_prepare_queries): their filter
        tables are built on the handle's side stream while the caller still computes / exchanges the probe lists.
        Returns the contiguous CUDA tensor to hand to that search; it must not be modified in between."""
        dev = self._device()
        h = self._sync_lists()
        xq = x.contiguous()
        if not (xq.is_cuda and xq.dtype == torch.float32 and xq.dim() == 2 and xq.shape[1] == self.d):
            raise AssertionError("prepare_queries takes a (nq, d) float32 CUDA tensor")
        with torch.cuda.device(dev):
            _lib.check(h.lib.b200_ivfpq_prepare_queries(h.h, xq.shape[0], xq.data_ptr(), _stream_ptr(dev)))
        return xq

    def search_preassigned_begin(self, x: torch.Tensor, k: int, list_ids: torch.Tensor, boot_lo: int, boot_hi: int):
        """First half of a sharded search (b200_ivfpq_search_preassigned_begin): returns the (nq,) int32 tensor of
        bootstrap-threshold bits (valid for the queries [boot_lo, boot_hi), +inf bits elsewhere), or None when the
        streaming pipeline does not apply to this shape."""
        dev = self._device()
        nq, nprobe = x.shape[0], int(list_ids.shape[1])
        self._check_search(k, nprobe)
        h = self._sync_lists()
        thr = torch.empty(nq, dtype=torch.int32, device=dev)
        self._split_keep = (x.contiguous(), list_ids.to(dev, torch.int64).contiguous())   # alive until _finish
        with torch.cuda.device(dev):
            rc = h.lib.b200_ivfpq_search_preassigned_begin(h.h, nq, self._split_keep[0].data_ptr(), k, nprobe,
                                                           self._split_keep[1].data_ptr(), int(boot_lo), int(boot_hi),
                                                           thr.data_ptr(), _stream_ptr(dev))
        if rc == 5:         # B200_IVFPQ_EUNSUPPORTED
            self._split_keep = None
            return None
        _lib.check(rc)
        return thr

    def search_preassigned_finish(self, thr: torch.Tensor, nq: int, k: int, out=None):
        dev = self._device()
        h = self._ensure_handle()
        if out is not None:
            D, I = _check_out(out, nq, k, dev)
        else:
            D = torch.empty((nq, k), dtype=torch.float32, device=dev)
            I = torch.empty((nq, k), dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(h.lib.b200_ivfpq_search_preassigned_finish(h.h, thr.data_ptr(), D.data_ptr(), I.data_ptr()))
        self._split_keep = None
        return D, I

    # ------------------------------------------------------------------ instrumentation (bench.py)
    def set_stage_timing(self, enable: bool):
        h = self._ensure_handle()
        _lib.check(h.lib.b200_ivfpq_set_stage_timing(h.h, 1 if enable else 0))

    def stage_ms(self):
        h = self._ensure_handle()
        out = (ctypes.c_float * 5)()
        _lib.check(h.lib.b200_ivfpq_get_stage_ms(h.h, out))
        return dict(zip(["coarse_dist", "coarse_select", "pair_setup", "scan", "merge"], [float(v) for v in out]))

    def last_scan_stats(self):
        h = self._ensure_handle()
        b, c = ctypes.c_int64(), ctypes.c_int64()
        _lib.check(h.lib.b200_ivfpq_get_last_scan_stats(h.h, ctypes.byref(b), ctypes.byref(c)))
        return {"bytes": int(b.value), "codes": int(c.value)}

    def filter_ms(self):
        """Device time of the streaming filter kernel in the last timed search, or None if it did not run."""
        h = self._ensure_handle()
        out = ctypes.c_float()
        if h.lib.b200_ivfpq_get_filter_ms(h.h, ctypes.byref(out)) != 0:
            return None
        return float(out.value)

    def filter_stats(self, reset: bool = True):
        """Counters of the per-query-table filter scan since the last reset (zeros unless the handle was created with
        B200_IVFPQ_QL_STATS=1): survivor entries, exact evaluations, work items."""
        h = self._ensure_handle()
        out = (ctypes.c_int64 * 3)()
        _lib.check(h.lib.b200_ivfpq_get_filter_stats(h.h, out, 1 if reset else 0))
        return {"survivor_entries": int(out[0]), "exact_evaluations": int(out[1]), "work_items": int(out[2])}

    # ------------------------------------------------------------------ flat views for the oracle / extraction
    def to_arrays(self):
        """The flat arrays the reference's extraction scripts dump: coarse (nlist, d), pq (M, 256, dsub),
        offsets (nlist+1), codes (ntotal, M), ids (ntotal)."""
        self._finalize_lists()
        dev = self._device()
        codes = self._codes if self._codes is not None else torch.zeros((0, self.pq.M), dtype=torch.uint8, device=dev)
        ids = self._ids if self._ids is not None else torch.zeros((0,), dtype=torch.int64, device=dev)
        return {
            "coarse": self.quantizer.xb_tensor().cpu().numpy(),
            "pq": self.pq._centroids.cpu().numpy(),
            "offsets": self._offsets.copy(),
            "codes": codes.cpu().numpy(),
            "ids": ids.cpu().numpy(),
        }
