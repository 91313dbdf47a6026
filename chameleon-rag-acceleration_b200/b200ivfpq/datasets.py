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
    """x = centre[c] + sigma * N(0, I) with ncentres Gaussian centres in [0, 1)^d.  Chunk i of a stream is
    generated from seed (stream_seed, i) so any rank can produce any chunk independently."""

    def __init__(self, d: int, ncentres: int, sigma: float = 0.08, device="cuda", seed: int = 7):
        self.d, self.ncentres, self.sigma, self.device = d, ncentres, sigma, device
        g = torch.Generator(device=device)
        g.manual_seed(seed)
        self.centres = torch.rand((ncentres, d), generator=g, device=device, dtype=torch.float32)

    def chunk(self, stream_seed: int, chunk_id: int, n: int) -> torch.Tensor:
        g = torch.Generator(device=self.device)
        g.manual_seed(stream_seed * 1000003 + chunk_id)
        which = torch.randint(0, self.ncentres, (n,), generator=g, device=self.device)
        x = torch.randn((n, self.d), generator=g, device=self.device, dtype=torch.float32)
        x.mul_(self.sigma).add_(self.centres[which])
        return x
