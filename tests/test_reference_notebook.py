"""The oracle against the reference's NORMATIVE restatement, run here: the functions of
Faiss_experiments/my_faiss_extract_scripts/IVFPQ_1B_search.ipynb cells 16 and 20 (`get_invlist`, `distance_full_vec`,
`construct_distance_table`, `estimate_distance(s)`, `search_single_query`, `search_batch_query`; SURVEY.md 8c file 1 --
the code the notebook shows equal to Faiss at :8094-8095).  The cells are read out of the .ipynb where it lies under
/root/reference and exec'd as they are (nothing is copied into the repo; the test is skipped where the reference is
not mounted); `faiss` inside them is this repo's package (only `rev_swig_ptr` is touched) and `index.invlists` is a
plain array-backed object with the accessors the notebook calls.

The notebook sums with numpy's pairwise float32 `np.sum` and accumulates table entries in Python floats, the contract
sequentially in fp32: ids must agree except for near ties, distances within north_star's 1e-5."""
import json
import os
import sys

import numpy as np
import pytest

import _util

NOTEBOOK = "/root/reference/Chameleon/Faiss_experiments/my_faiss_extract_scripts/IVFPQ_1B_search.ipynb"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not os.path.exists(NOTEBOOK), reason="the reference is only mounted in the build container")


class _InvLists:
    """What `index.invlists` gives the notebook: list_size / get_ids / get_codes / code_size over flat arrays."""

    def __init__(self, offsets, codes, ids):
        self.offsets, self.codes, self.ids = offsets, codes, ids
        self.code_size = codes.shape[1]

    def list_size(self, l):
        return int(self.offsets[l + 1] - self.offsets[l])

    def get_ids(self, l):
        return self.ids[self.offsets[l]:self.offsets[l + 1]]

    def get_codes(self, l):
        return self.codes[self.offsets[l]:self.offsets[l + 1]].reshape(-1)


def _notebook_namespace(invlists):
    sys.path.insert(0, os.path.join(ROOT, "chameleon-rag-acceleration_b200"))
    import b200ivfpq as faiss
    nb = json.load(open(NOTEBOOK))
    cells = ["".join(c["source"]) for c in nb["cells"] if c["cell_type"] == "code"]
    defs = [s for s in cells if "def get_invlist" in s or "def search_single_query" in s]
    assert len(defs) == 2, "notebook cells moved"

    class _Index:                                   # cell 16 starts with `invlists = index.invlists`
        pass

    idx = _Index()
    idx.invlists = invlists
    ns = {"np": np, "faiss": faiss, "index": idx}
    for src in defs:
        exec(compile(src, NOTEBOOK, "exec"), ns)
    return ns


@pytest.mark.parametrize("d,nlist,M,n,nq,nprobe,k", [(32, 8, 8, 600, 6, 3, 10), (64, 12, 16, 900, 4, 4, 5)])
def test_oracle_matches_the_notebooks_search(oracle, d, nlist, M, n, nq, nprobe, k):
    a = _util.make_index_arrays(oracle, 31, d, nlist, M, n)
    xq = _util.make_queries(8, a, nq)
    ns = _notebook_namespace(_InvLists(a["offsets"], a["codes"], a["ids"]))
    ids_nb, dist_nb = ns["search_batch_query"](xq, nprobe, k, a["coarse"], a["pq"])
    I_nb, D_nb = np.asarray(ids_nb, np.int64), np.asarray(dist_nb, np.float64)
    D, I = oracle.C.search(xq, a["coarse"], a["pq"], a["offsets"], a["codes"], a["ids"], nprobe, k)
    _util.assert_same_modulo_near_ties(D, I, D_nb.astype(np.float32), I_nb, 1e-5, "oracle vs notebook")
    assert (I == I_nb).mean() > 0.95


def test_oracle_lut_and_adc_match_the_notebooks_functions(oracle):
    a = _util.make_index_arrays(oracle, 5, 64, 4, 16, 300)
    ns = _notebook_namespace(_InvLists(a["offsets"], a["codes"], a["ids"]))
    q = _util.make_queries(1, a, 1)[0]
    res = (q - a["coarse"][2]).astype(np.float32)
    T_nb = ns["construct_distance_table"](res, a["pq"])
    T = oracle.C.lut(q, a["coarse"][2], a["pq"])
    np.testing.assert_allclose(T, T_nb, rtol=2e-6)
    codes = a["codes"][:50]
    d_nb = np.array([ns["estimate_distance"](c, T_nb) for c in codes])
    np.testing.assert_allclose(oracle.C.adc(T, codes), d_nb, rtol=2e-6)
    c0 = a["coarse"][1]
    assert abs(float(ns["distance_full_vec"](q, c0)) - float(oracle.C.l2sqr(q, c0))) <= 2e-6 * float(oracle.C.l2sqr(q, c0))
