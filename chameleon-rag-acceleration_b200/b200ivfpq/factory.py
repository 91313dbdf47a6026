"""index_factory and ParameterSpace -- the configuration surface of the reference's drivers.

  faiss.index_factory(d, "IVF4096,PQ16")            bench_cpu_performance.py:98
  key grammar (OPQ..,)?(IVF<nlist>),(PQ<m>|Flat)    bench_gpu_performance_OSDI.py:422-424
  ps = faiss.ParameterSpace(); ps.initialize(index); ps.set_index_parameters(index, "nprobe=32")
                                                    bench_cpu_performance.py:237-252
The north_star adds the explicit PQ<m>x<nbits> spelling; the reference only ever uses nbits = 8.
"""
from __future__ import annotations

import re

from .index import METRIC_L2, IndexFlatL2, IndexIVFPQ

_IVFPQ = re.compile(r"^IVF(\d+),PQ(\d+)(?:x(\d+))?$")
_OPQ = re.compile(r"^OPQ(\d+)(?:_(\d+))?,(.*)$")


def index_factory(d: int, description: str, metric: int = METRIC_L2):
    key = description.strip()
    if metric != METRIC_L2:
        raise RuntimeError("index_factory: only METRIC_L2 is supported")
    if key == "Flat":
        return IndexFlatL2(d)
    o = _OPQ.match(key)
    if o:
        # "OPQ16,IVF4096,PQ16" / "OPQ16_64,..." (bench_cpu_recall.py:54): rotation (+ reduction to 64 dims) in front
        from .transforms import IndexPreTransform, OPQMatrix
        vt = OPQMatrix(d, int(o.group(1)), int(o.group(2)) if o.group(2) else -1)
        sub = index_factory(vt.d_out, o.group(3), metric)
        if not isinstance(sub, IndexIVFPQ):
            raise RuntimeError(f"index_factory: '{key}': OPQ is supported in front of IVF<nlist>,PQ<m> only")
        if sub.pq.M != vt.M:
            raise RuntimeError(f"index_factory: '{key}': OPQ{vt.M} does not match PQ{sub.pq.M}")
        return IndexPreTransform(vt, sub)
    m = _IVFPQ.match(key)
    if not m:
        raise RuntimeError(f"index_factory: could not parse '{key}' (supported: 'Flat', 'IVF<nlist>,PQ<m>[x8]')")
    nlist, M = int(m.group(1)), int(m.group(2))
    nbits = int(m.group(3)) if m.group(3) else 8
    return IndexIVFPQ(IndexFlatL2(d), d, nlist, M, nbits)


class ParameterSpace:
    """Only `nprobe` exists on this path (bench_cpu_performance.py:252, bench_gpu_1bn.py GpuParameterSpace)."""

    def initialize(self, index):
        self._index = index

    @staticmethod
    def _set(index, name: str, value: float):
        if name != "nprobe":
            raise RuntimeError(f"ParameterSpace: could not set parameter {name}")
        index = getattr(index, "index", index)          # IndexPreTransform -> its IVF sub-index
        v = int(round(float(value)))
        if v < 1:
            raise RuntimeError("nprobe must be >= 1")
        index.nprobe = v

    def set_index_parameter(self, index, name: str, value: float):
        self._set(index, name, value)

    def set_index_parameters(self, index, params: str):
        for tok in params.split(","):
            tok = tok.strip()
            if not tok:
                continue
            if "=" not in tok:
                raise RuntimeError(f"ParameterSpace: could not parse '{tok}'")
            name, value = tok.split("=", 1)
            self._set(index, name.strip(), float(value))


GpuParameterSpace = ParameterSpace
