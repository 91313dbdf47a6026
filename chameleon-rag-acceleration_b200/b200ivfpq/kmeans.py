"""GPU k-means for index.train (build-time, not on the timed path).

Mirrors what the reference's drivers get from Faiss: Lloyd iterations (niter = 25), training set
subsampled to at most max_points_per_centroid points per centroid, empty clusters re-seeded by splitting
the largest one (call sites: bench_cpu_performance.py:98-109, bench_gpu_1bn.py:522-542, 583-594).
The assignment step is a plain library GEMM (torch.addmm); the centroid update is this repo's sequential segmented
sum (b200_ivfpq_segment_sums), so that training is reproducible bit for bit for fixed seeds: two bench runs -- or two
GPU counts -- then search the same index.  Parity does not depend on training: the oracle and the CUDA kernels are
always compared on the SAME trained codebooks.
"""
from __future__ import annotations

import torch


def _assign(x: torch.Tensor, c: torch.Tensor, chunk: int = 0):
    """argmin_c ||x - c||^2 via ||c||^2 - 2 x.c ; returns (labels i64, sum of min distances)."""
    if chunk <= 0:   # keep the (chunk, k) distance matrix around 1 GiB
        chunk = max(1024, min(1 << 18, (1 << 28) // max(1, c.shape[0])))
    cn = (c * c).sum(1)
    labels = torch.empty(x.shape[0], dtype=torch.int64, device=x.device)
    obj = 0.0
    for i0 in range(0, x.shape[0], chunk):
        xb = x[i0:i0 + chunk]
        dist = torch.addmm(cn.unsqueeze(0), xb, c.t(), alpha=-2.0)
        val, idx = dist.min(dim=1)
        labels[i0:i0 + chunk] = idx
        obj += float((val + (xb * xb).sum(1)).sum())
    return labels, obj


def _segment_sums(x: torch.Tensor, labels: torch.Tensor, k: int):
    """(sums (k, d), counts (k,)) of the rows of x grouped by label, reproducible bit for bit: on the GPU a stable sort
    of the labels + the library's sequential segmented sum (b200_ivfpq_segment_sums) instead of an atomic scatter-add."""
    counts = torch.bincount(labels, minlength=k)
    if not x.is_cuda:
        sums = torch.zeros((k, x.shape[1]), dtype=x.dtype)
        sums.index_add_(0, labels, x)
        return sums, counts
    from . import _lib
    lib = _lib.load()
    order = torch.argsort(labels, stable=True)
    start = torch.zeros(k + 1, dtype=torch.int64, device=x.device)
    start[1:] = torch.cumsum(counts, 0)
    sums = torch.empty((k, x.shape[1]), dtype=torch.float32, device=x.device)
    xc = x.contiguous()
    with torch.cuda.device(x.device):
        _lib.check(lib.b200_ivfpq_segment_sums(k, xc.shape[1], xc.data_ptr(), order.data_ptr(), start.data_ptr(),
                                               sums.data_ptr(), int(torch.cuda.current_stream(x.device).cuda_stream)))
    return sums, counts


def kmeans(x: torch.Tensor, k: int, niter: int = 25, seed: int = 1234, max_points_per_centroid: int = 256,
           verbose: bool = False) -> torch.Tensor:
    """x: (n, d) float32 on the GPU.  Returns (k, d) float32 centroids."""
    assert x.dim() == 2 and x.dtype == torch.float32
    n, d = x.shape
    if n < k:
        raise RuntimeError(f"Number of training points ({n}) should be at least as large as number of clusters ({k})")
    g = torch.Generator(device=x.device)
    g.manual_seed(seed)
    if n > k * max_points_per_centroid:
        perm = torch.randperm(n, generator=g, device=x.device)[:k * max_points_per_centroid]
        x = x[perm]
        n = x.shape[0]
    perm = torch.randperm(n, generator=g, device=x.device)[:k]
    c = x[perm].clone()
    for it in range(niter):
        labels, obj = _assign(x, c)
        sums, counts = _segment_sums(x, labels, k)
        nonempty = counts > 0
        c = torch.where(nonempty.unsqueeze(1), sums / counts.clamp(min=1).unsqueeze(1).to(sums.dtype), c)
        nempty = int((~nonempty).sum())
        if nempty:
            # split the largest clusters: empty centroid <- donor * (1 + eps), donor <- donor * (1 - eps)
            empty_idx = torch.nonzero(~nonempty).flatten()
            donors = torch.argsort(counts, descending=True)[:nempty]
            eps = 1.0 / 1024.0
            c[empty_idx] = c[donors] * (1.0 + eps)
            c[donors] = c[donors] * (1.0 - eps)
        if verbose:
            print(f"  kmeans iter {it}: objective {obj:.6g}, empty {nempty}")
    return c.contiguous()


def kmeans_subspaces(x: torch.Tensor, M: int, ksub: int = 256, niter: int = 25, seed: int = 1234,
                     init: torch.Tensor | None = None) -> torch.Tensor:
    """Per-subspace k-means for the product quantizer.  x: (n, d) residuals; returns (M, ksub, dsub).
    `init` (M, ksub, dsub) warm-starts the centroids (OPQ retrains the PQ a few iterations per rotation update)."""
    n, d = x.shape
    dsub = d // M
    xs = x.reshape(n, M, dsub).permute(1, 0, 2).contiguous()          # (M, n, dsub)
    g = torch.Generator(device=x.device)
    g.manual_seed(seed)
    perm = torch.randperm(n, generator=g, device=x.device)[:ksub]
    c = xs[:, perm, :].clone() if init is None else init.clone()        # (M, ksub, dsub)
    ar = torch.arange(M, device=x.device).unsqueeze(1)
    for _ in range(niter):
        cn = (c * c).sum(2)                                             # (M, ksub)
        dist = torch.baddbmm(cn.unsqueeze(1), xs, c.transpose(1, 2), alpha=-2.0)   # (M, n, ksub)
        labels = dist.argmin(dim=2)                                     # (M, n)
        flat = (labels + ar * ksub).reshape(-1)
        sums, counts = _segment_sums(xs.reshape(-1, dsub), flat, M * ksub)
        sums, counts = sums.reshape(M, ksub, dsub), counts.reshape(M, ksub)
        nonempty = counts > 0
        c = torch.where(nonempty.unsqueeze(2), sums / counts.clamp(min=1).unsqueeze(2).to(sums.dtype), c)
        if not bool(nonempty.all()):
            eps = 1.0 / 1024.0
            for m in range(M):
                empty_idx = torch.nonzero(~nonempty[m]).flatten()
                if empty_idx.numel():
                    donors = torch.argsort(counts[m], descending=True)[:empty_idx.numel()]
                    c[m, empty_idx] = c[m, donors] * (1.0 + eps)
                    c[m, donors] = c[m, donors] * (1.0 - eps)
    return c.contiguous()
