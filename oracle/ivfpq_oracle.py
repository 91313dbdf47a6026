"""CPU oracle for the IVF-PQ search path.  TEST INFRASTRUCTURE ONLY.

Only tests/, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
bench.py may import this module.  The product package (chameleon-rag-acceleration_b200/) never does.

Two implementations of the same arithmetic contract (see ivfpq_oracle.c header):

* ``C`` -- ctypes binding of ``libivfpq_oracle.so`` (C, OpenMP); used for anything sizeable and as the
  CPU baseline ("restated Faiss-CPU algorithm, not the Faiss binary").
* ``np_*`` -- a slow numpy twin, written directly from the reference's notebook functions
  (Chameleon/Faiss_experiments/my_faiss_extract_scripts/IVFPQ_1B_search.ipynb:7922-8028), vectorised only
  across independent outputs so that every accumulation keeps its sequential fp32 order.  It exists to
  cross-check the C file on small cases.

Parity status: LUT arithmetic pinned by the reference's literal KAT
(retrieval_accelerator/LUT_construction_PEs/LUT_construction_PE_D128_M32/src/host.cpp:44-109); coarse stage pinned
against the reference's own CPU cell selection run here (``ref_coarse``: vendored hnswlib brute force, host.cpp:516-581,
compiled into oracle/_ref/ from the headers under /root/reference); LUT and ADC arithmetic pinned bit for bit against
the reference's HLS kernels run as a C simulation (``ref_fpga_lut_adc``), and the whole search against the reference's
complete accelerator kernel (``ref_accel_search``); end-to-end search "parity unpinned" against the
Faiss binary (Faiss not installable here).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libivfpq_oracle.so")
FLT_MAX = np.float32(3.4028234663852886e38)

_f32p = ctypes.POINTER(ctypes.c_float)
_i64p = ctypes.POINTER(ctypes.c_int64)
_u8p = ctypes.POINTER(ctypes.c_uint8)


def build(force: bool = False) -> str:
    """Compile the C oracle (gcc, a few hundred ms).  Building the checker is not using it."""
    src = os.path.join(_HERE, "ivfpq_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _SO


# ---- oracle/_ref: code OF THE REFERENCE run here (row a1 only) -------------------------------------------------------
_REF_SO = os.path.join(_HERE, "_ref", "libref_coarse.so")
REF_SRC = "/root/reference/Chameleon/retrieval_accelerator/entire_accelerator_final_SIFT_M32/src"
_ref_lib = None


def build_ref(force: bool = False):
    """Compile oracle/_ref/: the reference's own coarse-quantizer code (its vendored hnswlib BruteforceSearch,
    host.cpp:516-581) behind ref_coarse_shim.cpp, and its HLS LUT-construction / ADC kernels as a C simulation behind
    ref_fpga_shim.cpp, from the sources where they lie under /root/reference.  Returns the library path, or None when the reference is not mounted and nothing was prebuilt
    (the GPU box only ever uses the prebuilt file)."""
    have_ref = os.path.exists(os.path.join(REF_SRC, "hnswlib", "hnswlib.h"))
    newest = max(os.path.getmtime(os.path.join(_HERE, f)) for f in
                 ("ref_coarse_shim.cpp", "ref_fpga_shim.cpp", "ref_accel_shim.cpp", "hls_csim/ap_int.h",
                  "hls_csim/hls_stream.h", "Makefile"))
    built = [_REF_SO] + [os.path.join(_HERE, "_ref", f"libref_{kind}_{v}.so") for v in FPGA_VARIANTS
                         for kind in ("fpga", "accel")]
    stale = any(not os.path.exists(b) or os.path.getmtime(b) < newest for b in built)
    if have_ref and (force or stale):
        subprocess.run(["make", "-C", _HERE, "-B", "ref"], check=True, capture_output=True)
    return _REF_SO if os.path.exists(_REF_SO) else None


def ref_coarse(xq, centroids, nprobe):
    """The reference's CPU cell selection (hnswlib brute force + searchKnn): (dis, ids), rows ascending.  Distances come
    from hnswlib's SIMD L2 (another summation order than the contract): equal to rounding, not bit for bit."""
    global _ref_lib
    if _ref_lib is None:
        path = build_ref()
        if path is None:
            raise FileNotFoundError("oracle/_ref/libref_coarse.so not built (reference not mounted)")
        _ref_lib = ctypes.CDLL(path)
        _ref_lib.ref_coarse_bruteforce.restype = ctypes.c_int
        _ref_lib.ref_coarse_bruteforce.argtypes = [ctypes.c_int64, ctypes.c_int, _f32p, ctypes.c_int64, _f32p,
                                                   ctypes.c_int, _i64p, _f32p]
        _ref_lib.ref_coarse_simd.restype = ctypes.c_char_p
    xq, centroids = _f32(xq), _f32(centroids)
    nq, d = xq.shape
    ids = np.empty((nq, nprobe), np.int64)
    dis = np.empty((nq, nprobe), np.float32)
    rc = _ref_lib.ref_coarse_bruteforce(centroids.shape[0], d, _p(centroids, _f32p), nq, _p(xq, _f32p), nprobe,
                                        _p(ids, _i64p), _p(dis, _f32p))
    if rc:
        raise RuntimeError(f"ref_coarse_bruteforce failed: {rc}")
    return dis, ids


FPGA_VARIANTS = {"SIFT_M16": (128, 16), "SIFT_M32": (128, 32), "Deep_M16": (96, 16), "Deep_M32": (96, 32)}
_fpga_libs = {}


def _fpga_lib(variant):
    D, M = FPGA_VARIANTS[variant]
    if variant not in _fpga_libs:
        if build_ref() is None or not os.path.exists(os.path.join(_HERE, "_ref", f"libref_fpga_{variant}.so")):
            raise FileNotFoundError(f"oracle/_ref/libref_fpga_{variant}.so not built (reference not mounted)")
        lib = ctypes.CDLL(os.path.join(_HERE, "_ref", f"libref_fpga_{variant}.so"))
        d_, m_ = ctypes.c_int(), ctypes.c_int()
        lib.ref_fpga_dims(ctypes.byref(d_), ctypes.byref(m_))
        assert (d_.value, m_.value) == (D, M), "constants.hpp of the variant changed"
        lib.ref_fpga_lut_adc.restype = ctypes.c_int
        lib.ref_fpga_lut_adc.argtypes = [ctypes.c_int, ctypes.c_int, _f32p, _f32p, _f32p, ctypes.POINTER(ctypes.c_int),
                                         _u8p, _f32p, _f32p]
        lib.ref_fpga_queue_l1.restype = ctypes.c_int
        lib.ref_fpga_queue_l1.argtypes = [ctypes.c_int, ctypes.c_int, _f32p, ctypes.POINTER(ctypes.c_int), _f32p]
        _fpga_libs[variant] = lib
    return _fpga_libs[variant]


def ref_fpga_queue(dist, k, variant="SIFT_M16"):
    """The reference's systolic priority queue (priority_queue_L1.hpp, queue length k in {1, 10, 100}) fed the
    candidates `dist` in scan order.  Returns (offsets, distances) of the occupied slots sorted by (distance, offset)."""
    dist = _f32(dist)
    off = np.empty(k, np.int32)
    od = np.empty(k, np.float32)
    rc = _fpga_lib(variant).ref_fpga_queue_l1(k, dist.shape[0], _p(dist, _f32p),
                                              off.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), _p(od, _f32p))
    if rc:
        raise ValueError(f"queue length {k} not instantiated (1, 10, 100)")
    keep = od < np.float32(9e9)                      # LARGE_NUM marks an empty slot
    off, od = off[keep], od[keep]
    order = np.lexsort((off, od))
    return off[order].astype(np.int64), od[order]


_accel_libs = {}


def ref_accel_dims(variant):
    """(D, M, TOPK, ADC_PE_NUM, PRIORITY_QUEUE_LEN_L1) the accelerator build was compiled with (its constants.hpp)."""
    if variant not in _accel_libs:
        path = os.path.join(_HERE, "_ref", f"libref_accel_{variant}.so")
        if build_ref() is None or not os.path.exists(path):
            raise FileNotFoundError(f"oracle/_ref/libref_accel_{variant}.so not built (reference not mounted)")
        lib = ctypes.CDLL(path)
        lib.ref_accel_run.restype = None
        lib.ref_accel_run.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p] * 11
        _accel_libs[variant] = lib
    v = [ctypes.c_int() for _ in range(5)]
    _accel_libs[variant].ref_accel_dims(*[ctypes.byref(x) for x in v])
    return tuple(x.value for x in v)


def ref_accel_search(variant, centroids, pq, offsets, codes, ids, xq, probes):
    """The reference's COMPLETE accelerator kernel (`vadd`, retrieval_accelerator/entire_accelerator_final_<variant>/src/
    vadd.cpp) run as a C simulation on the probed cells `probes` (nq, nprobe): returns (D, I) of shape (nq, TOPK) sorted
    by (distance, id); TOPK = 100 is a compile-time constant of the reference.

    This function is the HOST SIDE, written here from the formats the kernel parses (the reference's host.cpp needs the
    Xilinx runtime and is not used):
      * meta_data_init (vadd.cpp load_meta_data / DRAM_utils.hpp:185-222): per cell the PQ-code start word, the vector-id
        start index, the cell size; then the PQ codebook (M, 256, D/M) as raw floats;
      * PQ codes (DRAM_utils.hpp:103-183, load_PQ_codes): four banks of 512-bit words; vector j of a cell is entry
        j // ADC_PE_NUM of PE s = j % ADC_PE_NUM, i.e. bank s // lanes, lane a = s % lanes (lanes = ADC_PE_NUM / 4),
        bytes [a*M, (a+1)*M) of that bank's word `start + entry`;
      * vector ids (hierarchical_priority_queue.hpp:185-263): bank b, index `vid_start + entry * lanes + a`;
      * query packets (vadd.cpp:38-97, network_input_processing): header word (query id, nprobe), cell ids, the query,
        the nprobe centroids, all in 512-bit words, zero padded;
      * results (vadd.cpp:99-160): header word, TOPK 64-bit ids, TOPK fp32 distances.
    The accelerator is approximate by design (its first-level queues keep PRIORITY_QUEUE_LEN_L1 entries each); callers
    pick inputs on which that truncation does not bite, and every first-level queue must receive at least that many
    candidates per query (empty queue slots carry uninitialised cell ids in the reference's code)."""
    D, M, TOPK, PE, L1 = ref_accel_dims(variant)
    lib = _accel_libs[variant]
    centroids, pq, xq = _f32(centroids), _f32(pq), _f32(xq)
    codes, ids, offsets = _u8(codes).reshape(-1, M), _i64(ids), _i64(offsets)
    probes = np.ascontiguousarray(probes, np.int64)
    assert pq.shape == (M, 256, D // M) and centroids.shape[1] == D and xq.shape[1] == D
    lanes = PE // 4
    nlist = offsets.shape[0] - 1
    sizes = np.diff(offsets)
    words = (sizes + PE - 1) // PE
    pq_start = np.zeros(nlist, np.int64)
    pq_start[1:] = np.cumsum(words)[:-1]
    vid_start = pq_start * lanes
    nq, nprobe = probes.shape
    per_query = sizes[probes].sum(axis=1)
    if per_query.min() < 4 * PE * L1:
        raise ValueError("every query must scan enough vectors to fill all first-level queues of the accelerator")
    total_words = max(int(words.sum()), 1)
    bank_codes = np.zeros((4, total_words, 64), np.uint8)
    bank_ids = np.zeros((4, total_words * lanes), np.uint64)
    list_no = np.repeat(np.arange(nlist), sizes)
    j = np.arange(codes.shape[0]) - offsets[list_no]              # position inside the cell
    entry, s = j // PE, j % PE
    bank, lane = s // lanes, s % lanes
    word = pq_start[list_no] + entry
    bank_codes.reshape(4, total_words, 64 // M, M)[bank, word, lane] = codes
    bank_ids[bank, vid_start[list_no] + entry * lanes + lane] = ids.astype(np.uint64)
    meta = np.ascontiguousarray(np.concatenate([pq_start.astype(np.int32), vid_start.astype(np.int32),
                                                sizes.astype(np.int32), pq.reshape(-1).view(np.int32)]))
    wvec = (D * 4 + 63) // 64
    wcell = (nprobe * 4 + 63) // 64
    packets = np.zeros((nq, 1 + wcell + wvec + nprobe * wvec, 16), np.int32)
    packets[:, 0, 0] = np.arange(nq)
    packets[:, 0, 1] = nprobe
    cells = np.zeros((nq, wcell * 16), np.int32)
    cells[:, :nprobe] = probes
    packets[:, 1:1 + wcell] = cells.reshape(nq, wcell, 16)
    vec = np.zeros((nq, wvec * 16), np.float32)
    vec[:, :D] = xq
    packets[:, 1 + wcell:1 + wcell + wvec] = vec.view(np.int32).reshape(nq, wvec, 16)
    cen = np.zeros((nq, nprobe, wvec * 16), np.float32)
    cen[:, :, :D] = centroids[probes]
    packets[:, 1 + wcell + wvec:] = cen.view(np.int32).reshape(nq, nprobe * wvec, 16)
    wid, wdist = (TOPK * 8 + 63) // 64, (TOPK * 4 + 63) // 64
    out = np.zeros((nq, 1 + wid + wdist, 64), np.uint8)
    lib.ref_accel_run(nq, nlist, nprobe, meta.ctypes.data, packets.ctypes.data,
                      *[bank_codes[b].ctypes.data for b in range(4)], *[bank_ids[b].ctypes.data for b in range(4)],
                      out.ctypes.data)
    I = np.empty((nq, TOPK), np.int64)
    Dist = np.empty((nq, TOPK), np.float32)
    for q in range(nq):
        iq = out[q, 1:1 + wid].reshape(-1).view(np.uint64)[:TOPK].astype(np.int64)
        dq = out[q, 1 + wid:].reshape(-1).view(np.float32)[:TOPK]
        order = np.lexsort((iq, dq))
        I[q], Dist[q] = iq[order], dq[order]
    return Dist, I


def ref_fpga_lut_adc(variant, pq, xq, centers, nscan, codes):
    """The reference's OWN HLS kernels (LUT_construction.hpp: LUT_construction_wrapper; ADC.hpp: PQ_lookup_computation of
    retrieval_accelerator/entire_accelerator_final_<variant>/src) run as a C simulation (ref_fpga_shim.cpp).
    pq (M, 256, dsub); xq (nq, D); centers (nq, nprobe, D) = the probed cells' centroids; nscan (nq, nprobe) entries
    scanned per cell; codes (sum nscan, M) in (query, probe, entry) order.
    Returns (lut (nq, nprobe, M, 256), dist (sum nscan,))."""
    D, M = FPGA_VARIANTS[variant]
    lib = _fpga_lib(variant)
    pq, xq, centers = _f32(pq), _f32(xq), _f32(centers)
    nscan = np.ascontiguousarray(nscan, np.int32)
    codes = _u8(codes).reshape(-1, M)
    nq, nprobe = nscan.shape
    assert pq.shape == (M, 256, D // M) and xq.shape == (nq, D) and centers.shape == (nq, nprobe, D)
    assert codes.shape[0] == int(nscan.sum())
    lut = np.empty((nq, nprobe, 256, M), np.float32)
    dist = np.empty(codes.shape[0], np.float32)
    rc = lib.ref_fpga_lut_adc(nq, nprobe, _p(pq, _f32p), _p(xq, _f32p), _p(centers, _f32p),
                              nscan.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), _p(codes, _u8p), _p(lut, _f32p),
                              _p(dist, _f32p))
    if rc:
        raise RuntimeError(f"ref_fpga_lut_adc failed: {rc}")
    return np.ascontiguousarray(lut.transpose(0, 1, 3, 2)), dist


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _p(a, t):
    return a.ctypes.data_as(t)


class _C:
    """ctypes facade over libivfpq_oracle.so."""

    def __init__(self):
        self._lib = None

    @property
    def lib(self):
        if self._lib is None:
            if not os.path.exists(_SO):
                build()
            lib = ctypes.CDLL(_SO)
            lib.oracle_l2sqr.restype = ctypes.c_float
            lib.oracle_l2sqr.argtypes = [_f32p, _f32p, ctypes.c_int]
            lib.oracle_num_threads.restype = ctypes.c_int
            lib.oracle_set_num_threads.argtypes = [ctypes.c_int]
            lib.oracle_coarse.argtypes = [ctypes.c_int64, ctypes.c_int, _f32p, ctypes.c_int64, _f32p, ctypes.c_int,
                                          _i64p, _f32p]
            lib.oracle_lut.argtypes = [ctypes.c_int, ctypes.c_int, _f32p, _f32p, _f32p, _f32p]
            lib.oracle_adc.argtypes = [ctypes.c_int, _f32p, ctypes.c_int64, _u8p, _f32p]
            lib.oracle_search_preassigned.argtypes = [ctypes.c_int64, ctypes.c_int, _f32p, ctypes.c_int64, _f32p,
                                                      ctypes.c_int, _f32p, _i64p, _u8p, _i64p, ctypes.c_int, _i64p,
                                                      ctypes.c_int, _f32p, _i64p]
            lib.oracle_search.argtypes = [ctypes.c_int64, ctypes.c_int, _f32p, ctypes.c_int64, _f32p, ctypes.c_int,
                                          _f32p, _i64p, _u8p, _i64p, ctypes.c_int, ctypes.c_int, _f32p, _i64p, _i64p,
                                          _f32p]
            lib.oracle_assign.argtypes = [ctypes.c_int64, ctypes.c_int, _f32p, ctypes.c_int64, _f32p, _i64p]
            lib.oracle_encode.argtypes = [ctypes.c_int64, ctypes.c_int, _f32p, _f32p, _i64p, ctypes.c_int, _f32p, _u8p]
            lib.oracle_merge_shards.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_int, _f32p, _i64p, _f32p, _i64p]
            self._lib = lib
        return self._lib

    def num_threads(self) -> int:
        return int(self.lib.oracle_num_threads())

    def set_num_threads(self, n: int) -> None:
        self.lib.oracle_set_num_threads(int(n))

    def l2sqr(self, a, b) -> np.float32:
        a, b = _f32(a), _f32(b)
        return np.float32(self.lib.oracle_l2sqr(_p(a, _f32p), _p(b, _f32p), a.shape[0]))

    def coarse(self, xq, centroids, nprobe):
        xq, centroids = _f32(xq), _f32(centroids)
        nq, d = xq.shape
        ids = np.empty((nq, nprobe), np.int64)
        dis = np.empty((nq, nprobe), np.float32)
        self.lib.oracle_coarse(nq, d, _p(xq, _f32p), centroids.shape[0], _p(centroids, _f32p), nprobe, _p(ids, _i64p),
                               _p(dis, _f32p))
        return dis, ids

    def lut(self, q, c, pq):
        q, c, pq = _f32(q), _f32(c), _f32(pq)
        M = pq.shape[0]
        T = np.empty((M, 256), np.float32)
        self.lib.oracle_lut(q.shape[0], M, _p(q, _f32p), _p(c, _f32p), _p(pq, _f32p), _p(T, _f32p))
        return T

    def adc(self, T, codes):
        T, codes = _f32(T), _u8(codes)
        n, M = codes.shape
        out = np.empty(n, np.float32)
        self.lib.oracle_adc(M, _p(T, _f32p), n, _p(codes, _u8p), _p(out, _f32p))
        return out

    def search_preassigned(self, xq, centroids, pq, offsets, codes, ids, probe_ids, k):
        xq, centroids, pq = _f32(xq), _f32(centroids), _f32(pq)
        offsets, codes, ids, probe_ids = _i64(offsets), _u8(codes), _i64(ids), _i64(probe_ids)
        nq, d = xq.shape
        D = np.empty((nq, k), np.float32)
        I = np.empty((nq, k), np.int64)
        self.lib.oracle_search_preassigned(nq, d, _p(xq, _f32p), centroids.shape[0], _p(centroids, _f32p), pq.shape[0],
                                           _p(pq, _f32p), _p(offsets, _i64p), _p(codes, _u8p), _p(ids, _i64p),
                                           probe_ids.shape[1], _p(probe_ids, _i64p), k, _p(D, _f32p), _p(I, _i64p))
        return D, I

    def search(self, xq, centroids, pq, offsets, codes, ids, nprobe, k, return_probes=False):
        xq, centroids, pq = _f32(xq), _f32(centroids), _f32(pq)
        offsets, codes, ids = _i64(offsets), _u8(codes), _i64(ids)
        nq, d = xq.shape
        D = np.empty((nq, k), np.float32)
        I = np.empty((nq, k), np.int64)
        pid = np.empty((nq, nprobe), np.int64)
        pdis = np.empty((nq, nprobe), np.float32)
        self.lib.oracle_search(nq, d, _p(xq, _f32p), centroids.shape[0], _p(centroids, _f32p), pq.shape[0],
                               _p(pq, _f32p), _p(offsets, _i64p), _p(codes, _u8p), _p(ids, _i64p), nprobe, k,
                               _p(D, _f32p), _p(I, _i64p), _p(pid, _i64p), _p(pdis, _f32p))
        if return_probes:
            return D, I, pdis, pid
        return D, I

    def assign(self, x, centroids):
        x, centroids = _f32(x), _f32(centroids)
        out = np.empty(x.shape[0], np.int64)
        self.lib.oracle_assign(x.shape[0], x.shape[1], _p(x, _f32p), centroids.shape[0], _p(centroids, _f32p),
                               _p(out, _i64p))
        return out

    def encode(self, x, centroids, list_no, pq):
        x, centroids, pq, list_no = _f32(x), _f32(centroids), _f32(pq), _i64(list_no)
        M = pq.shape[0]
        codes = np.empty((x.shape[0], M), np.uint8)
        self.lib.oracle_encode(x.shape[0], x.shape[1], _p(x, _f32p), _p(centroids, _f32p), _p(list_no, _i64p), M,
                               _p(pq, _f32p), _p(codes, _u8p))
        return codes

    def merge_shards(self, Ds, Is):
        """Ds, Is: (nshard, nq, k).  bench_multi_cpu_performance_OSDI.py:203-219."""
        Ds, Is = _f32(Ds), _i64(Is)
        nshard, nq, k = Ds.shape
        D = np.empty((nq, k), np.float32)
        I = np.empty((nq, k), np.int64)
        self.lib.oracle_merge_shards(nshard, nq, k, _p(Ds, _f32p), _p(Is, _i64p), _p(D, _f32p), _p(I, _i64p))
        return D, I


C = _C()


# --------------------------------------------------------------------------------------------------
# numpy twin (small cases only)
# --------------------------------------------------------------------------------------------------

def np_l2sqr_rows(a, B):
    """L2^2 from vector a to every row of B; sequential fp32 accumulation over j (ipynb:7922-7927)."""
    a, B = _f32(a), _f32(B)
    acc = np.zeros(B.shape[0], np.float32)
    for j in range(B.shape[1]):
        diff = (a[j] - B[:, j]).astype(np.float32)
        acc = (acc + (diff * diff).astype(np.float32)).astype(np.float32)
    return acc


def np_coarse(xq, centroids, nprobe):
    """ipynb:7991-7999: distance to every centroid, sort (ties -> lower id), take nprobe."""
    xq = _f32(xq)
    ids = np.empty((xq.shape[0], nprobe), np.int64)
    dis = np.empty((xq.shape[0], nprobe), np.float32)
    for q in range(xq.shape[0]):
        dd = np_l2sqr_rows(xq[q], centroids)
        order = np.lexsort((np.arange(dd.shape[0]), dd))[:nprobe]
        ids[q], dis[q] = order, dd[order]
    return dis, ids


def np_lut(q, c, pq):
    """construct_distance_table (ipynb:7929-7946) on the residual q - c (ipynb:8006)."""
    q, c, pq = _f32(q), _f32(c), _f32(pq)
    M, ksub, dsub = pq.shape
    res = (q - c).astype(np.float32)
    T = np.zeros((M, ksub), np.float32)
    for j in range(dsub):
        diff = (res.reshape(M, dsub)[:, j][:, None] - pq[:, :, j]).astype(np.float32)
        T = (T + (diff * diff).astype(np.float32)).astype(np.float32)
    return T


def np_adc(T, codes):
    """estimate_distance (ipynb:7948-7960): sum over m ascending, fp32."""
    T, codes = _f32(T), _u8(codes)
    acc = np.zeros(codes.shape[0], np.float32)
    for m in range(codes.shape[1]):
        acc = (acc + T[m, codes[:, m]]).astype(np.float32)
    return acc


def np_search_preassigned(xq, centroids, pq, offsets, codes, ids, probe_ids, k):
    """search_single_query (ipynb:8001-8017) with the (distance, scan order) total order."""
    xq = _f32(xq)
    nq = xq.shape[0]
    D = np.full((nq, k), FLT_MAX, np.float32)
    I = np.full((nq, k), -1, np.int64)
    for q in range(nq):
        dd, ii = [], []
        for l in probe_ids[q]:
            if l < 0:
                continue
            beg, end = int(offsets[l]), int(offsets[l + 1])
            if end == beg:
                continue
            T = np_lut(xq[q], centroids[l], pq)
            dd.append(np_adc(T, codes[beg:end]))
            ii.append(ids[beg:end])
        if not dd:
            continue
        dd, ii = np.concatenate(dd), np.concatenate(ii)
        order = np.lexsort((np.arange(dd.shape[0]), dd))[:k]
        D[q, :order.shape[0]] = dd[order]
        I[q, :order.shape[0]] = ii[order]
    return D, I


def np_search(xq, centroids, pq, offsets, codes, ids, nprobe, k):
    _, pid = np_coarse(xq, centroids, nprobe)
    return np_search_preassigned(xq, centroids, pq, offsets, codes, ids, pid, k)


def np_merge_shards(Ds, Is):
    """bench_multi_cpu_performance_OSDI.py:203-219: concatenate, stable argsort, take k."""
    nshard, nq, k = Ds.shape
    D = np.full((nq, k), FLT_MAX, np.float32)
    I = np.full((nq, k), -1, np.int64)
    for q in range(nq):
        dd, ii = Ds[:, q, :].reshape(-1), Is[:, q, :].reshape(-1)
        keep = ii >= 0
        dd, ii = dd[keep], ii[keep]
        order = np.argsort(dd, kind="stable")[:k]
        D[q, :order.shape[0]] = dd[order]
        I[q, :order.shape[0]] = ii[order]
    return D, I


# --------------------------------------------------------------------------------------------------
# recall (bench_gpu_performance_OSDI.py:196-200, 690-692)
# --------------------------------------------------------------------------------------------------

def recall_at_k(I, gt, k):
    total = 0
    for gt_row, row in zip(gt[:, :k], I[:, :k]):
        total += np.intersect1d(gt_row, row).shape[0]
    return total / float(gt[:, :k].size)


def r1_at_k(I, gt, k):
    return float((I[:, :k] == gt[:, :1]).sum()) / I.shape[0]


# ---------------------------------------------------------------------------------------------------------
# Faiss-CPU's own distance form (use_precomputed_table = 1), restated from upstream faiss 1.7.x IndexIVFPQ.cpp
# (precompute_table / IVFPQScannerT::precompute_list_tables_L2); the source is NOT under /root/reference.
# It is NOT the parity contract (that is the residual-LUT form above, the notebook's); it exists to measure how far
# the two forms are apart -- the reason BASELINE states a 1e-5 relative tolerance instead of bit-exactness.
# ---------------------------------------------------------------------------------------------------------
def np_precomputed_table(centroids, pq):
    """P[l][m][c] = ||p_mc||^2 + 2 <c_l[m], p_mc>   (nlist, M, ksub) fp32."""
    centroids, pq = _f32(centroids), _f32(pq)
    M, ksub, dsub = pq.shape
    cm = centroids.reshape(centroids.shape[0], M, dsub)
    norms = np.einsum("mkj,mkj->mk", pq, pq).astype(np.float32)
    ip = np.einsum("lmj,mkj->lmk", cm, pq).astype(np.float32)
    return (norms[None] + np.float32(2.0) * ip).astype(np.float32)


def np_search_preassigned_precomputed(xq, centroids, pq, offsets, codes, ids, probe_ids, k, table=None):
    """dis = ||q - c_l||^2 + sum_m (P[l][m][code_m] - 2 <q_m, p_m,code_m>), fp32, sum over m ascending from dis0."""
    xq, centroids, pq = _f32(xq), _f32(centroids), _f32(pq)
    M, ksub, dsub = pq.shape
    P = np_precomputed_table(centroids, pq) if table is None else table
    nq = xq.shape[0]
    D = np.full((nq, k), FLT_MAX, np.float32)
    I = np.full((nq, k), -1, np.int64)
    for q in range(nq):
        qip = np.einsum("mj,mkj->mk", xq[q].reshape(M, dsub), pq).astype(np.float32)      # <q_m, p_mc>
        dd, ii = [], []
        for l in probe_ids[q]:
            if l < 0:
                continue
            beg, end = int(offsets[l]), int(offsets[l + 1])
            if end == beg:
                continue
            T = (P[l] - np.float32(2.0) * qip).astype(np.float32)
            acc = np.full(end - beg, np_l2sqr_rows(xq[q], centroids[l:l + 1])[0], np.float32)
            cc = codes[beg:end]
            for m in range(M):
                acc = (acc + T[m, cc[:, m]]).astype(np.float32)
            dd.append(acc)
            ii.append(ids[beg:end])
        if not dd:
            continue
        dd, ii = np.concatenate(dd), np.concatenate(ii)
        order = np.lexsort((np.arange(dd.shape[0]), dd))[:k]
        D[q, :order.shape[0]] = dd[order]
        I[q, :order.shape[0]] = ii[order]
    return D, I
