"""The oracle against every golden vector the reference holds for this path."""
import os

import numpy as np

import _util


def _kat():
    return np.load(os.path.join(_util.GOLDEN, "lut_kat_d128_m32.npz"))


def test_lut_kat_c_oracle_bit_exact(oracle):
    """LUT_construction_PE_D128_M32/src/host.cpp:44-109: literal query / centroid / pq[i] = i % 256."""
    z = _kat()
    T = oracle.C.lut(z["query"], z["centroid"], z["pq"])
    _util.assert_bit_equal(T, z["lut"], "C oracle LUT vs reference KAT")


def test_lut_kat_numpy_twin_bit_exact(oracle):
    z = _kat()
    _util.assert_bit_equal(oracle.np_lut(z["query"], z["centroid"], z["pq"]), z["lut"], "numpy twin LUT vs KAT")


def test_lut_kat_reference_tolerance_rule(oracle):
    """The reference's own pass rule (host.cpp:176-191): |hw - sw| <= 0.01 and 0.99 <= hw/sw <= 1.01."""
    z = _kat()
    hw, sw = oracle.C.lut(z["query"], z["centroid"], z["pq"]).astype(np.float64), z["lut"].astype(np.float64)
    assert (np.abs(hw - sw) <= 0.01).all()
    nz = sw != 0
    assert ((hw[nz] / sw[nz] <= 1.01) & (hw[nz] / sw[nz] >= 0.99)).all()


def test_lut_kat_layout_is_m_major():
    """host.cpp:47-55 reshapes (M, 256, D/M) -> (256, D): golden pq must follow the (M, 256, dsub) layout the
    extraction scripts document (extract_Enzian_U250_required_data.py:222-232)."""
    z = _kat()
    assert z["pq"].shape == (32, 256, 4)
    flat = z["pq"].reshape(-1)
    assert (flat == (np.arange(flat.size) % 256)).all()


def test_reference_coarse_fixture_conventions(oracle):
    """node0_info/np_cell_{ids,dists}_batch.npy (Faiss output for 32 SIFT1B queries, nprobe 32, IVF8192): the
    centroids are not in the tree, so only the output conventions can be pinned: ascending distances, distinct
    in-range ids.  The oracle's coarse output must follow the same conventions."""
    ids = np.load(os.path.join(_util.GOLDEN, "ref_np_cell_ids_batch.npy"))
    dis = np.load(os.path.join(_util.GOLDEN, "ref_np_cell_dists_batch.npy"))
    assert ids.shape == dis.shape == (32, 32)
    assert (np.diff(dis, axis=1) >= 0).all()
    assert ids.min() >= 0 and ids.max() < 8192
    assert all(len(set(r.tolist())) == 32 for r in ids)
    rng = np.random.default_rng(5)
    cent = rng.random((200, 16), dtype=np.float32)
    xq = rng.random((32, 16), dtype=np.float32)
    odis, oids = oracle.C.coarse(xq, cent, 32)
    assert (np.diff(odis, axis=1) >= 0).all()
    assert oids.min() >= 0 and oids.max() < 200
    assert all(len(set(r.tolist())) == 32 for r in oids)
    assert oids.dtype == np.int64 and odis.dtype == np.float32


def test_search_small_regression(oracle):
    """Frozen end-to-end vector (make_search_golden.py): C oracle and numpy twin both reproduce it."""
    z = np.load(os.path.join(_util.GOLDEN, "search_small.npz"))
    nprobe, k = int(z["nprobe"]), int(z["k"])
    D, I, pdis, pid = oracle.C.search(z["xq"], z["coarse"], z["pq"], z["offsets"], z["codes"], z["ids"], nprobe, k,
                                      return_probes=True)
    _util.assert_bit_equal(D, z["D"], "D")
    _util.assert_bit_equal(I, z["I"], "I")
    _util.assert_bit_equal(pid, z["probe_ids"], "probe ids")
    _util.assert_bit_equal(pdis, z["probe_dis"], "probe distances")
    D2, I2 = oracle.np_search(z["xq"], z["coarse"], z["pq"], z["offsets"], z["codes"], z["ids"], nprobe, k)
    _util.assert_bit_equal(D2, z["D"], "numpy twin D")
    _util.assert_bit_equal(I2, z["I"], "numpy twin I")
