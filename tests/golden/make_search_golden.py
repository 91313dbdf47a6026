"""Freeze the oracle's end-to-end output on a small seeded index -> tests/golden/search_small.npz.
Regenerate only when the arithmetic contract changes on purpose."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ivfpq_oracle as oracle   # noqa: E402
import _util                                 # noqa: E402

if __name__ == "__main__":
    oracle.build()
    a = _util.make_index_arrays(oracle, seed=20261018, d=32, nlist=24, M=8, n=3000, used_lists=20)
    xq = _util.make_queries(77, a, 16)
    D, I, pdis, pid = oracle.C.search(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], nprobe=6, k=10,
                                      return_probes=True)
    np.savez_compressed(os.path.join(HERE, "search_small.npz"), coarse=a["coarse"], pq=a["pq"], offsets=a["offsets"],
                        codes=a["codes"], ids=a["ids"], xq=xq, D=D, I=I, probe_ids=pid, probe_dis=pdis,
                        nprobe=6, k=10)
    print("wrote search_small.npz", D[0, :3], I[0, :3])
