"""Synthetic data generators for the BASELINE.json configs.

  uniform  -- the reference's own toy generator (Faiss_experiments/IVFPQ_random_dataset.py:6-13):
              np.random.seed(1234); x = random((n, d)).astype('float32'); x[:, 0] += arange(n) / 1000.
  clustered -- mixture of Gaussians (the reference warns that uniform data gives very low PQ recall,
              generate_SYN_dataset.py:3-4); generated chunk-wise on the GPU, deterministic per chunk, so that
              100M-1B vector bases never exist in full.
Seeds: base 1234, queries 4321, train 999.
"""
from __future__ import annotations

import numpy as np
import torch

SEED_BASE, SEED_QUERY, SEED_TRAIN = 1234, 4321, 999


def uniform_reference(nb: int, nq: int, d: int):
    """Bit-for-bit the arrays of IVFPQ_random_dataset.py:6-13."""
    np.random.seed(1234)
    xb = np.random.random((nb, d)).astype("float32")
    xb[:, 0] += np.arange(nb) / 1000.0
    xq = np.random.random((nq, d)).astype("float32")
    xq[:, 0] += np.arange(nq) / 1000.0
    return xb, xq


class ClusteredGenerator:
    """x = centre[c] + sigma * (z @ B) + sigma_iso * N(0, I_d), z ~ N(0, I_r): ncentres Gaussian blobs in [0, 1)^d
    whose spread lives on an r-dimensional subspace (latent_dim = r; 0 means isotropic in all d dimensions).
    Real descriptors (SIFT, Deep) have an intrinsic dimension of 10-20, which is what makes PQ recall meaningful;
    isotropic 128-d noise makes all cluster members equidistant and PQ recall collapses.
    Chunk i of a stream is generated from seed (stream_seed, i), so any rank can produce any chunk independently."""

    def __init__(self, d: int, ncentres: int, sigma: float = 0.1, device="cuda", seed: int = 7, latent_dim: int = 12,
                 sigma_iso: float = 0.002):
        self.d, self.ncentres, self.sigma, self.device = d, ncentres, sigma, device
        self.latent_dim, self.sigma_iso = latent_dim, sigma_iso
        g = torch.Generator(device=device)
        g.manual_seed(seed)
        self.centres = torch.rand((ncentres, d), generator=g, device=device, dtype=torch.float32)
        if latent_dim > 0:
            self.basis = torch.randn((latent_dim, d), generator=g, device=device, dtype=torch.float32)
            self.basis /= self.basis.norm(dim=1, keepdim=True)

    def chunk(self, stream_seed: int, chunk_id: int, n: int) -> torch.Tensor:
        g = torch.Generator(device=self.device)
        g.manual_seed(stream_seed * 1000003 + chunk_id)
        which = torch.randint(0, self.ncentres, (n,), generator=g, device=self.device)
        if self.latent_dim > 0:
            z = torch.randn((n, self.latent_dim), generator=g, device=self.device, dtype=torch.float32)
            x = torch.addmm(self.centres[which], z, self.basis, alpha=self.sigma)
            if self.sigma_iso > 0:
                x.add_(torch.randn((n, self.d), generator=g, device=self.device, dtype=torch.float32), alpha=self.sigma_iso)
            return x
        x = torch.randn((n, self.d), generator=g, device=self.device, dtype=torch.float32)
        x.mul_(self.sigma).add_(self.centres[which])
        return x
