"""OPQ pre-transform and IndexPreTransform.

The reference builds half of its recall studies on keys of the form "OPQ16,IVF4096,PQ16"
(Faiss_experiments/README.md:67-92, bench_cpu_recall.py:54, train_SYN_dataset.py:24) and reads the result back as
    linear_trans = faiss.downcast_VectorTransform(index.chain.at(0));  OPQ_mat = vector_to_array(linear_trans.A)
    downcasted_index = faiss.downcast_index(index.index)
(my_faiss_extract_scripts/extract_FPGA_required_data.py:162-172).  Those names are kept.

The transform is a d_in x d_out rotation applied to every vector before the IVF-PQ index sees it; training is the
non-parametric OPQ alternation (Ge et al., as in Faiss OPQMatrix::train): rotate, (re)train the PQ for a few Lloyd
iterations, reconstruct, solve the orthogonal Procrustes problem with an SVD.  It is build-time work and a plain
library GEMM / SVD (torch); the search hot path is unchanged -- the rotated queries go through the same kernels, and
parity with the oracle is checked on the rotated inputs.
"""
from __future__ import annotations

import numpy as np
import torch

from .index import IndexIVFPQ, _as_f32_matrix, _require_cuda, _to_device
from .kmeans import kmeans_subspaces


class VectorTransform:
    """faiss.VectorTransform: what the drivers call a pre-processor (bench_gpu_performance_OSDI.py:230, d_in / d_out /
    apply_py); OPQMatrix is the one the reference's factory strings produce."""
    d_in = 0
    d_out = 0
    is_trained = False

    def apply_py(self, x):
        raise NotImplementedError


class OPQMatrix(VectorTransform):
    """y = A x with A (d_out, d_in) row-major, the layout of Faiss LinearTransform.A; b = 0."""

    def __init__(self, d: int, M: int, d2: int = -1):
        self.d_in, self.d_out, self.M = int(d), int(d2 if d2 > 0 else d), int(M)
        if self.d_out % self.M:
            raise RuntimeError(f"OPQMatrix: d_out = {self.d_out} is not a multiple of M = {self.M}")
        if self.d_out > self.d_in:
            raise RuntimeError("OPQMatrix: d_out > d_in")
        self.niter, self.niter_pq, self.niter_pq_0 = 50, 4, 40        # Faiss defaults
        self.max_train_points = 256 * 256
        self.have_bias = False
        self.is_trained = False
        self.verbose = False
        self._A: torch.Tensor | None = None                           # (d_out, d_in) on the GPU

    # Faiss-style accessors (extract_FPGA_required_data.py:165-167 reshapes A to (d, d))
    @property
    def A(self) -> np.ndarray:
        return np.zeros(0, np.float32) if self._A is None else self._A.reshape(-1).cpu().numpy()

    @property
    def b(self) -> np.ndarray:
        return np.zeros(0, np.float32)

    def set_matrix(self, A):
        A = torch.as_tensor(np.ascontiguousarray(A, np.float32) if not isinstance(A, torch.Tensor) else A)
        self._A = A.reshape(self.d_out, self.d_in).to(_require_cuda()).float().contiguous()
        self.is_trained = True

    def train(self, x):
        x = _as_f32_matrix(x, self.d_in, "OPQMatrix.train")
        dev = _require_cuda()
        xt = _to_device(x, dev)
        g = torch.Generator(device=dev)
        g.manual_seed(1234)
        if xt.shape[0] > self.max_train_points:
            xt = xt[torch.randperm(xt.shape[0], generator=g, device=dev)[:self.max_train_points]]
        n = xt.shape[0]
        if n < 256:
            raise RuntimeError(f"OPQMatrix.train: {n} training points, need at least 256")
        # random orthonormal start (d_in, d_out)
        q, _ = torch.linalg.qr(torch.randn((self.d_in, self.d_in), generator=g, device=dev))
        R = q[:, :self.d_out].contiguous()
        pq = None
        dsub = self.d_out // self.M
        ar = torch.arange(self.M, device=dev).unsqueeze(1)
        for it in range(self.niter):
            xr = xt @ R                                                  # (n, d_out)
            pq = kmeans_subspaces(xr, self.M, 256, niter=self.niter_pq_0 if it == 0 else self.niter_pq, seed=4321, init=pq)
            xs = xr.reshape(n, self.M, dsub).permute(1, 0, 2)             # (M, n, dsub)
            cn = (pq * pq).sum(2)
            labels = torch.baddbmm(cn.unsqueeze(1), xs, pq.transpose(1, 2), alpha=-2.0).argmin(dim=2)   # (M, n)
            y = pq.reshape(self.M * 256, dsub)[(labels + ar * 256).reshape(-1)].reshape(self.M, n, dsub)
            y = y.permute(1, 0, 2).reshape(n, self.d_out)                 # reconstruction in the rotated space
            if self.verbose:
                print(f"  OPQ iter {it}: quantisation error {float(((xr - y) ** 2).sum() / n):.6g}")
            # orthogonal Procrustes: R = argmin ||x R - y||_F  s.t.  R^T R = I
            U, _, Vt = torch.linalg.svd(xt.t() @ y, full_matrices=False)  # (d_in, d_out), (d_out, d_out)
            R = (U @ Vt).contiguous()
        self._A = R.t().contiguous()
        self.is_trained = True

    def apply(self, x):
        """numpy in -> numpy out, torch in -> torch out (Faiss: apply_py)."""
        x = _as_f32_matrix(x, self.d_in, "OPQMatrix.apply")
        if not self.is_trained:
            raise RuntimeError("Error: 'is_trained' failed (OPQMatrix.apply before train)")
        y = _to_device(x, self._A.device) @ self._A.t()
        return y if isinstance(x, torch.Tensor) else y.cpu().numpy()

    apply_py = apply


class _Chain:
    def __init__(self, items):
        self._items = list(items)

    def at(self, i):
        return self._items[i]

    def size(self):
        return len(self._items)


class IndexPreTransform:
    """faiss.IndexPreTransform(vt, sub_index): every vector goes through the transform chain first."""

    def __init__(self, vt: OPQMatrix, index: IndexIVFPQ):
        if vt.d_out != index.d:
            raise RuntimeError(f"IndexPreTransform: transform outputs {vt.d_out} dimensions, index expects {index.d}")
        self.chain = _Chain([vt])
        self.index = index
        self.d = vt.d_in
        self.verbose = False

    # ---- state forwarded to the sub-index (ParameterSpace sets nprobe on the outer index: bench_cpu_recall.py)
    @property
    def nprobe(self):
        return self.index.nprobe

    @nprobe.setter
    def nprobe(self, v):
        self.index.nprobe = v

    @property
    def ntotal(self):
        return self.index.ntotal

    @property
    def is_trained(self):
        return self.chain.at(0).is_trained and self.index.is_trained

    @property
    def parallel_mode(self):
        return self.index.parallel_mode

    @parallel_mode.setter
    def parallel_mode(self, v):
        self.index.parallel_mode = v

    def _transform(self, x, what):
        x = _as_f32_matrix(x, self.d, what)
        return self.chain.at(0).apply(_to_device(x, _require_cuda())), isinstance(x, torch.Tensor)

    def train(self, x):
        x = _as_f32_matrix(x, self.d, "train")
        vt = self.chain.at(0)
        if not vt.is_trained:
            vt.train(x)
        self.index.train(vt.apply(_to_device(x, _require_cuda())))

    def add(self, x):
        y, _ = self._transform(x, "add")
        self.index.add(y)

    def add_with_ids(self, x, ids):
        y, _ = self._transform(x, "add")
        self.index.add_with_ids(y, ids)

    def search(self, x, k: int):
        y, is_torch = self._transform(x, "search")
        D, I = self.index.search(y, k)
        return (D, I) if is_torch else (D.cpu().numpy(), I.cpu().numpy())

    def search_preassigned(self, x, k: int, list_ids):
        y, is_torch = self._transform(x, "search_preassigned")
        D, I = self.index.search_preassigned(y, k, list_ids)
        if is_torch or isinstance(D, np.ndarray):
            return D, I
        return D.cpu().numpy(), I.cpu().numpy()


def downcast_VectorTransform(vt):
    return vt


def write_VectorTransform(vt: OPQMatrix, fname: str) -> None:
    """faiss.write_VectorTransform(preproc, cachefile) (bench_gpu_1bn.py:507): the trained matrix as one .npz."""
    if not isinstance(vt, OPQMatrix) or not vt.is_trained:
        raise RuntimeError("write_VectorTransform: a trained OPQMatrix is required")
    with open(fname, "wb") as f:
        np.savez(f, kind="OPQMatrix", d_in=vt.d_in, d_out=vt.d_out, M=vt.M, A=vt.A.reshape(vt.d_out, vt.d_in))


def read_VectorTransform(fname: str) -> OPQMatrix:
    """faiss.read_VectorTransform(cachefile) (bench_gpu_performance_OSDI.py:520, bench_gpu_1bn.py:510)."""
    z = np.load(fname)
    if str(z["kind"]) != "OPQMatrix":
        raise RuntimeError(f"read_VectorTransform: unsupported transform {z['kind']}")
    vt = OPQMatrix(int(z["d_in"]), int(z["M"]), int(z["d_out"]))
    vt.set_matrix(z["A"])
    return vt
