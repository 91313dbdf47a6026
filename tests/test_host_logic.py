"""Host-side logic that needs no GPU: factory grammar, ParameterSpace, shard arithmetic, the C-ABI exports,
and that the product path fails loudly (never falls back) without a CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import _util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol():
    import b200ivfpq
    header = open(os.path.join(ROOT, "include", "b200_ivfpq.h")).read()
    declared = re.findall(r"B200_API\s+[\w\s\*]+?\b(b200_ivfpq_\w+)\s*\(", header)
    assert len(declared) >= 16
    lib = ctypes.CDLL(b200ivfpq.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/b200_ivfpq.h but not exported"
    from b200ivfpq._lib import SYMBOLS
    assert sorted(declared) == sorted(s[0] for s in SYMBOLS), "ctypes table out of sync with the header"
    assert b"sm_100a" in b200ivfpq.load_library().b200_ivfpq_version()


def test_cabi_argument_errors_without_gpu():
    """Pure argument validation happens before any CUDA call."""
    import b200ivfpq
    lib = b200ivfpq.load_library()
    h = ctypes.c_void_p()
    assert lib.b200_ivfpq_create(128, 16, 16, 4, ctypes.byref(h)) == 5          # nbits != 8 -> EUNSUPPORTED
    assert b"nbits" in lib.b200_ivfpq_last_error()
    assert lib.b200_ivfpq_create(100, 16, 16, 8, ctypes.byref(h)) == 1          # d % m != 0 -> EINVAL
    assert lib.b200_ivfpq_create(0, 16, 16, 8, ctypes.byref(h)) == 1
    assert lib.b200_ivfpq_merge_shards(0, 1, 1, None, None, None, None, None) == 1
    assert lib.b200_ivfpq_launch_count() >= 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    import b200ivfpq as faiss
    lib = faiss.load_library()
    h = ctypes.c_void_p()
    rc = lib.b200_ivfpq_create(128, 16, 16, 8, ctypes.byref(h))
    assert rc == 3 and b"no CPU path" in lib.b200_ivfpq_last_error()
    index = faiss.index_factory(32, "IVF4,PQ4")
    with pytest.raises(RuntimeError, match="no CPU path"):
        index.train(np.zeros((100, 32), np.float32))
    with pytest.raises(RuntimeError):
        faiss.IndexFlatL2(8).add(np.zeros((4, 8), np.float32))


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing in the product package may import, link, load or execute it
    (comments that cite the contract are fine)."""
    pkg = os.path.join(ROOT, "chameleon-rag-acceleration_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            path = os.path.join(dp, f)
            if f.endswith(".py"):
                text = open(path).read()
                for pat in ("import oracle", "from oracle", "ivfpq_oracle", "libivfpq_oracle"):
                    assert pat not in text, (path, pat)
            elif f.endswith((".cu", ".cuh", ".h", ".sh")):
                for line in open(path):
                    if "oracle" in line.lower():
                        code = line.split("//")[0]
                        assert "oracle" not in code.lower(), (path, line)
                        assert "#include" not in line and "dlopen" not in line, (path, line)


def test_index_factory_grammar():
    import b200ivfpq as faiss
    for key, nlist, m in [("IVF1024,PQ16", 1024, 16), ("IVF8192,PQ16x8", 8192, 16), ("IVF65536,PQ16x8", 65536, 16),
                          ("IVF16384,PQ64x8", 16384, 64), ("IVF8192,PQ32x8", 8192, 32)]:
        d = 768 if m == 64 else 128
        idx = faiss.index_factory(d, key)
        assert isinstance(idx, faiss.IndexIVFPQ)
        assert (idx.nlist, idx.pq.M, idx.pq.nbits, idx.d, idx.nprobe, idx.ntotal) == (nlist, m, 8, d, 1, 0)
        assert idx.pq.dsub == d // m and idx.pq.ksub == 256 and idx.pq.code_size == m
        assert not idx.is_trained
    assert isinstance(faiss.index_factory(64, "Flat"), faiss.IndexFlatL2)
    opq = faiss.index_factory(128, "OPQ16,IVF1024,PQ16")          # bench_cpu_recall.py:54
    assert isinstance(opq, faiss.IndexPreTransform) and isinstance(opq.index, faiss.IndexIVFPQ)
    assert (opq.d, opq.index.d, opq.chain.size(), opq.chain.at(0).d_out, opq.is_trained) == (128, 128, 1, 128, False)
    opq64 = faiss.index_factory(128, "OPQ16_64,IVF256,PQ16")
    assert (opq64.d, opq64.index.d, opq64.index.pq.dsub) == (128, 64, 4)
    ps = faiss.ParameterSpace()
    ps.set_index_parameters(opq, "nprobe=12")
    assert opq.nprobe == 12 and opq.index.nprobe == 12
    for bad in ["IVF1024,PQ16x4", "OPQ8,IVF1024,PQ16", "OPQ16,IVF1024,Flat", "IVF1024,Flat", "IMI2x8,PQ16", "IVF,PQ16",
                "garbage"]:
        with pytest.raises(RuntimeError):
            faiss.index_factory(128, bad)
    with pytest.raises(RuntimeError):
        faiss.index_factory(100, "IVF16,PQ16")    # d % M


def test_parameter_space():
    import b200ivfpq as faiss
    idx = faiss.index_factory(128, "IVF1024,PQ16")
    ps = faiss.ParameterSpace()
    ps.initialize(idx)
    ps.set_index_parameters(idx, "nprobe=32")        # bench_cpu_performance.py:252
    assert idx.nprobe == 32
    ps.set_index_parameter(idx, "nprobe", 4.0)
    assert idx.nprobe == 4
    with pytest.raises(RuntimeError):
        ps.set_index_parameters(idx, "ht=64")
    with pytest.raises(RuntimeError):
        ps.set_index_parameters(idx, "nprobe")
    idx.parallel_mode = 3                             # faiss_retriever.py:71, accepted and ignored
    faiss.omp_set_num_threads(4)


def test_shard_positions_partition():
    from b200ivfpq.shards import shard_positions
    for n, world in [(10, 3), (1, 4), (0, 2), (1000, 8)]:
        parts = [shard_positions(n, r, world) for r in range(world)]
        allp = np.sort(np.concatenate(parts))
        assert np.array_equal(allp, np.arange(n))
        assert all((p % world == r).all() for r, p in enumerate(parts))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_pack_unpack_roundtrip_cpu():
    from b200ivfpq.shards import pack_results, unpack_results
    g = torch.Generator().manual_seed(0)
    D = torch.rand((5, 7), generator=g)
    I = torch.randint(-1, 2 ** 40, (5, 7), generator=g, dtype=torch.int64)
    buf = pack_results(D, I)
    assert buf.dtype == torch.int32 and buf.shape == (5, 7, 3)
    D2, I2 = unpack_results(buf)
    assert torch.equal(D2, D) and torch.equal(I2, I)


def test_uniform_reference_generator_matches_the_script():
    """datasets.uniform_reference restates IVFPQ_random_dataset.py:6-13."""
    from b200ivfpq.datasets import uniform_reference
    xb, xq = uniform_reference(1000, 10, 16)
    np.random.seed(1234)
    ref = np.random.random((1000, 16)).astype("float32")
    ref[:, 0] += np.arange(1000) / 1000.
    assert np.array_equal(xb, ref) and xq.shape == (10, 16) and xq.dtype == np.float32


def test_faiss_container_roundtrip(oracle, tmp_path):
    """IwPQ container (Faiss 1.7.x layout): writer -> parser round trip, both flat-storage variants, full and sparse
    list-size tables.  Not a check against a Faiss binary (none is available): see faiss_io.py STATUS."""
    from b200ivfpq.faiss_io import parse_faiss_ivfpq, write_faiss_ivfpq
    for used, storage in [(None, "float"), (3, "bytes")]:
        a = _util.make_index_arrays(oracle, 5, 32, 12, 8, 700, used_lists=used)
        fn = os.path.join(tmp_path, f"toy_IVF12,PQ8_populated_{storage}.index")
        write_faiss_ivfpq(fn, a, nprobe=7, flat_storage=storage)
        raw = open(fn, "rb").read()
        assert raw[:4] == b"IwPQ" and b"IxF2" in raw[:64] and b"ilar" in raw
        assert (b"sprs" in raw) == (used is not None)
        z = parse_faiss_ivfpq(fn)
        assert (z["d"], z["nlist"], z["M"], z["nbits"], z["nprobe"]) == (32, 12, 8, 8, 7)
        for key in ("coarse", "pq", "offsets", "codes", "ids"):
            _util.assert_bit_equal(z[key], a[key], key)
    with pytest.raises(RuntimeError):
        parse_faiss_ivfpq(b"IxF2" + bytes(64))


def test_faiss_container_bytes_by_hand(tmp_path):
    """The coarse quantizer's flat storage, byte for byte, built by hand from the Faiss write macros as documented in
    faiss_io.py: the default writer emits count = nlist * d followed by the raw float32 bytes (WRITEVECTOR /
    WRITEXBVECTOR on the uint8 image, i.e. what Faiss <= 1.6 and >= 1.7.2 read); the optional 'float' variant emits
    count = nlist * d / 4 (WRITEXBVECTOR on the float vector); the reader takes both."""
    import struct
    from b200ivfpq.faiss_io import parse_faiss_ivfpq, write_faiss_ivfpq
    nlist, d, M = 4, 8, 2
    coarse = np.arange(nlist * d, dtype=np.float32).reshape(nlist, d)
    pq = (np.arange(M * 256 * (d // M), dtype=np.float32) * 0.5).reshape(M, 256, d // M)
    offsets = np.array([0, 2, 2, 3, 3], np.int64)
    codes = np.array([[1, 2], [3, 4], [5, 6]], np.uint8)
    ids = np.array([10, 11, 12], np.int64)
    a = {"coarse": coarse, "pq": pq, "offsets": offsets, "codes": codes, "ids": ids}

    def by_hand(count):
        hdr = lambda dd, n: struct.pack("<iqqq?i", dd, n, 1 << 20, 1 << 20, True, 1)
        out = b"IwPQ" + hdr(d, 3) + struct.pack("<QQ", nlist, 5)
        out += b"IxF2" + hdr(d, nlist) + struct.pack("<Q", count) + coarse.tobytes()
        out += struct.pack("<bQ", 0, 0) + struct.pack("<?Q", True, M)
        out += struct.pack("<QQQQ", d, M, 8, pq.size) + pq.tobytes()
        out += b"ilar" + struct.pack("<QQ", nlist, M)
        out += b"sprs" + struct.pack("<Q", 4) + struct.pack("<QQQQ", 0, 2, 2, 1)
        out += codes[0:2].tobytes() + ids[0:2].tobytes() + codes[2:3].tobytes() + ids[2:3].tobytes()
        return out

    for storage, count in (("bytes", nlist * d), ("float", nlist * d // 4)):
        fn = os.path.join(tmp_path, f"hand_{storage}.index")
        write_faiss_ivfpq(fn, a, nprobe=5, flat_storage=storage)
        assert open(fn, "rb").read() == by_hand(count), storage
        z = parse_faiss_ivfpq(by_hand(count))
        for key in ("coarse", "pq", "offsets", "codes", "ids"):
            _util.assert_bit_equal(z[key], a[key], key)
    # the default is the count every current Faiss release reads
    fn = os.path.join(tmp_path, "default.index")
    write_faiss_ivfpq(fn, a, nprobe=5)
    assert open(fn, "rb").read() == by_hand(nlist * d)


def test_sass_is_blackwell_native_and_exact():
    """Static evidence from the built library (cuobjdump, no GPU needed): the coarse GEMM uses tcgen05 / TMEM / TMA,
    the two-query scan uses LDS.64 + FFMA2, and no kernel on the exact-arithmetic path contains a contracted
    multiply-add it did not ask for (ptxas fuses mul.rn.f32x2 + add.rn.f32x2 into FFMA2: see scan_duo.cuh)."""
    import shutil
    import subprocess
    import b200ivfpq
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", b200ivfpq.LIB_PATH], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            funcs[cur].append(line)

    def ops(substr):
        names = [n for n in funcs if substr in n]
        assert names, f"no kernel matching {substr}"
        return names, "\n".join("\n".join(funcs[n]) for n in names)

    _, gemm = ops("coarse_tc_gemm_kernel")
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR"):
        assert mnemonic in gemm, f"{mnemonic} missing from the coarse GEMM"
    names, _ = ops("scan_duo16_kernelILi8E")
    duo = "\n".join(funcs[names[0]])
    assert duo.count("LDS.64") >= 16 * 7 and duo.count("FFMA2") >= 2 * 16 * 7      # 7 unrolled 16-step blocks
    assert "FMUL2" not in duo or duo.count("FFMA2") % 32 == 0
    # exact kernels never use a scalar FFMA that the source did not write (__fmaf_rn appears only in the one-query
    # scan's 0/1-multiplier accumulate); the LUT / distance arithmetic must stay FADD / FMUL / FADD2 / FMUL2
    # (coarse_rescore_kernel is left out: its sqrtf for the error bound legitimately expands to FFMA)
    for kern in ("coarse_dist_kernel", "encode_kernel", "coarse_exact_flagged_kernel"):
        _, text = ops(kern)
        assert not re.search(r"\bFFMA\b", text), f"contracted FFMA in {kern}"
    # round 2's filter kernels (csrc/scan_stream.cuh).  Four queries: 16 LDS.64 look-ups per code and unrolled block,
    # the table copy with 128-bit loads / stores, no fp32 multiply on the code path besides the one FFMA per
    # (code, query) of the threshold test, no local memory.  Two queries: the look-ups are 32-bit, the codes arrive by
    # LDGSTS (cp.async) and come back with one LDS.128 per code.  The bulk-async experiment keeps UBLKCP + SYNCS.
    four = "\n".join(funcs[[n for n in funcs if "st_filter_kernelILi16ELb0E" in n][0]])
    assert four.count("LDS.64") >= 16 * 4 and four.count("STS.128") >= 8 and "LDGSTS" in four   # (group descriptor)
    assert not re.search(r"\b(LDL|STL)\b", four), "the filter kernel spills"
    two = "\n".join(funcs[[n for n in funcs if "st_filter_kernelILi16ELb1E" in n][0]])
    assert len(re.findall(r"\bLDS R", two)) >= 16 * 2, "32-bit look-ups"
    assert two.count("LDS.64") < 16 and two.count("LDGSTS.E.BYPASS.128") >= 7 + 2 and two.count("LDS.128") >= 2
    assert not re.search(r"\b(LDL|STL)\b", two), "the two-query filter kernel spills"
    _, bulk = ops("st_filter2_kernel")
    assert "UBLKCP" in bulk and "SYNCS" in bulk
    evals = "\n".join(funcs[[n for n in funcs if "st_eval_kernelILi16ELb1E" in n][0]])
    assert evals.count("LDS.128") >= 2, "codebook rows gathered from shared memory"


def test_faiss_named_shims():
    """Names the reference's drivers call around the hot path."""
    import b200ivfpq as faiss
    res = faiss.StandardGpuResources()
    res.setTempMemory(1 << 20)
    vres, vdev = faiss.GpuResourcesVector(), faiss.IntVector()              # bench_gpu_performance_OSDI.py:492-503
    vres.push_back(res)
    vdev.push_back(0)
    assert vres.size() == 1 and vdev.at(0) == 0
    co = faiss.GpuMultipleClonerOptions()
    co.shard, co.useFloat16, co.usePrecomputed, co.indicesOptions = True, False, True, 0      # :587-595
    idx = faiss.index_factory(64, "IVF16,PQ8")
    # single process: the index is already on the GPU, cloning is the identity (no process group -> no sharding)
    assert faiss.index_cpu_to_gpu_multiple(vres, vdev, idx, co) is idx
    assert faiss.index_cpu_to_gpus_list(idx, co=co, gpus=[0]) is idx and faiss.index_gpu_to_cpu(idx) is idx
    a = np.arange(10, dtype=np.int64)
    assert np.array_equal(faiss.rev_swig_ptr(faiss.swig_ptr(a), 4), a[:4])
    faiss.cvar.indexIVF_stats.reset()
    assert faiss.cvar.indexIVF_stats.nq == 0 and faiss.cvar.indexIVFPQ_stats.search_cycles == 0
    # `from faiss.contrib.ivf_tools import search_preassigned` (faiss_server.py:24): forwards to the index method
    from b200ivfpq.contrib.ivf_tools import search_preassigned

    class _Probe:
        def search_preassigned(self, xq, k, list_nos):
            return ("called", k, list_nos)

    assert search_preassigned(_Probe(), None, 7, [[1, -1]], coarse_dis=None) == ("called", 7, [[1, -1]])


REF_ROOT = "/root/reference/Chameleon"
# the drivers of the hot path (SURVEY.md section 8b "What calls it") and of its build / evaluation tooling (8a a10, 8f)
REF_DRIVERS = ["llm_inference_gpu/ralm/server/faiss_server.py", "llm_inference_gpu/ralm/retriever/faiss_retriever.py",
               "llm_inference_gpu/ralm/index_scanner/index_scanner.py", "Faiss_experiments/bench_cpu_performance.py",
               "Faiss_experiments/bench_gpu_performance_OSDI.py", "Faiss_experiments/IVFPQ_random_dataset.py",
               "Faiss_experiments/bench_multi_cpu_performance_OSDI.py", "Faiss_experiments/bench_gpu_1bn.py"]


@pytest.mark.skipif(not os.path.exists(REF_ROOT), reason="the reference is only mounted in the build container")
def test_every_faiss_name_the_references_drivers_use_exists():
    """`import b200ivfpq as faiss` must resolve every `faiss.<name>` the reference's drivers of this path mention
    (comment lines aside): a user switching over finds the surface they call."""
    import re
    import b200ivfpq as faiss
    missing = {}
    for f in REF_DRIVERS:
        code = "\\n".join(ln for ln in open(os.path.join(REF_ROOT, f)).read().splitlines()
                         if not ln.strip().startswith("#"))
        names = set(re.findall(r"\\bfaiss\\.([A-Za-z_][A-Za-z0-9_]*)", code))
        gone = sorted(n for n in names if not hasattr(faiss, n))
        if gone:
            missing[f] = gone
    assert not missing, missing
    from b200ivfpq.contrib.ivf_tools import search_preassigned        # faiss_server.py:24, faiss_retriever.py:14
    assert callable(search_preassigned)


def test_ground_truth_heap_and_helpers():
    """float_maxheap_array_t as bench_gpu_1bn.py:427-456 drives it, ranklist_intersection_size (:222), read_index's
    second argument (bench_cpu_performance.py:116)."""
    import inspect
    import b200ivfpq as faiss
    rng = np.random.default_rng(0)
    nq, k, nb = 7, 5, 60
    allD = rng.random((nq, nb)).astype(np.float32)
    allD[:, 10] = allD[:, 3]                                           # an exact tie between two database entries
    gt_I = np.zeros((nq, k), dtype="int64")
    gt_D = np.zeros((nq, k), dtype="float32")
    heaps = faiss.float_maxheap_array_t()
    heaps.k, heaps.nh = k, nq
    heaps.val, heaps.ids = faiss.swig_ptr(gt_D), faiss.swig_ptr(gt_I)
    heaps.heapify()
    assert (gt_I == -1).all()
    for i0 in range(0, nb, 20):                                        # blocks of the database, as compute_GT() does
        blk = allD[:, i0:i0 + 20]
        I = np.argsort(blk, axis=1, kind="stable")[:, :k]
        D = np.take_along_axis(blk, I, axis=1)
        I = I + i0
        heaps.addn_with_ids(k, faiss.swig_ptr(D), faiss.swig_ptr(I.astype("int64")), k)
    heaps.reorder()
    want = np.argsort(allD, axis=1, kind="stable")[:, :k]
    assert np.array_equal(gt_I, want) and np.array_equal(gt_D, np.take_along_axis(allD, want, axis=1))
    assert faiss.ranklist_intersection_size(3, faiss.swig_ptr(np.array([1, 2, 3, 4])), 3,
                                            faiss.swig_ptr(np.array([3, 9, 1, 2]))) == 2
    assert list(inspect.signature(faiss.read_index).parameters)[:2] == ["fname", "io_flags"] and faiss.IO_FLAG_MMAP
    # the manual replica container of bench_gpu_performance_OSDI.py:613-626, in one process
    class _Member:
        d, ntotal, nprobe = 8, 123, 4
        this = faiss.IndexFlatL2.this

        def search(self, x, k):
            return "D", "I"

    rep = faiss.IndexReplicas()
    for _ in range(2):
        m = _Member()
        m.this.disown()
        rep.addIndex(m)
    rep.own_fields = True
    import torch
    assert rep.count() == 2 and rep.ntotal == 123 and rep.d == 8 and rep.search(torch.zeros((3, 8)), 5) == ("D", "I")
    faiss.ParameterSpace().set_index_parameter(rep, "nprobe", 9)
    assert rep.at(0).nprobe == 9
    for unsupported in (faiss.IndexIVFFlat, faiss.PCAMatrix):
        with pytest.raises(RuntimeError):
            unsupported(None, 8, 4)


def test_clustering_has_no_cpu_path():
    import torch
    import b200ivfpq as faiss
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError, match="no CPU path"):
        faiss.Clustering(6, 4).train(np.zeros((64, 6), np.float32))


@pytest.mark.gpu
def test_clustering_wrapper_trains_centroids():
    """faiss.Clustering(d, k).train(x, index) as bench_gpu_1bn.py:520-542 calls it."""
    import torch
    import b200ivfpq as faiss
    rng = np.random.default_rng(1)
    centres = rng.random((8, 6)).astype(np.float32) * 10
    x = (centres[rng.integers(0, 8, 4000)] + rng.standard_normal((4000, 6)).astype(np.float32) * 0.05).astype(np.float32)

    class _Sink:
        def reset(self):
            self.xb = None

        def add(self, c):
            self.xb = np.asarray(c.cpu() if isinstance(c, torch.Tensor) else c)

    clus = faiss.Clustering(6, 16)          # twice as many centroids as blobs: Lloyd then covers every blob
    clus.niter = 20
    clus.max_points_per_centroid = 10000000
    sink = _Sink()
    clus.train(x, sink)
    c = faiss.vector_float_to_array(clus.centroids).reshape(16, 6)
    assert sink.xb.shape == (16, 6) and np.allclose(sink.xb, c)
    # every true centre has a learned centroid next to it
    d2 = ((centres[:, None, :] - c[None, :, :]) ** 2).sum(2)
    assert (d2.min(axis=1) < 0.05).all(), d2.min(axis=1)


@pytest.mark.gpu
def test_training_is_reproducible():
    """Same training set and seeds -> the same codebooks bit for bit (the centroid update is a sequential segmented sum,
    not an atomic scatter-add): bench lines of different runs and GPU counts come from the same index."""
    import torch
    import b200ivfpq as faiss
    g = torch.Generator(device="cuda")
    g.manual_seed(3)
    x = torch.randn((60000, 32), generator=g, device="cuda")
    books = []
    for _ in range(2):
        index = faiss.index_factory(32, "IVF256,PQ8")
        index.cp_niter = 6
        index.train(x)
        books.append((index.quantizer.xb_tensor().clone(), index.pq.centroids_tensor().clone()))
    assert torch.equal(books[0][0], books[1][0]) and torch.equal(books[0][1], books[1][1])


def test_vector_transform_persistence(tmp_path, monkeypatch):
    """faiss.write_VectorTransform / read_VectorTransform (bench_gpu_1bn.py:507-510) on the OPQ matrix; the device
    placement is patched out so that the file format round-trips without a GPU."""
    import torch
    import b200ivfpq as faiss
    from b200ivfpq import transforms
    monkeypatch.setattr(transforms, "_require_cuda", lambda: torch.device("cpu"))
    rng = np.random.default_rng(2)
    A = np.linalg.qr(rng.standard_normal((12, 12)))[0][:8].astype(np.float32)       # (d_out = 8, d_in = 12)
    vt = faiss.OPQMatrix(12, 4, 8)
    vt.set_matrix(A)
    assert isinstance(vt, faiss.VectorTransform) and vt.is_trained and (vt.d_in, vt.d_out) == (12, 8)
    fn = str(tmp_path / "opq.vt")
    faiss.write_VectorTransform(vt, fn)
    vt2 = faiss.read_VectorTransform(fn)
    assert (vt2.d_in, vt2.d_out, vt2.M) == (12, 8, 4) and np.array_equal(vt2.A, vt.A)
    x = rng.standard_normal((5, 12)).astype(np.float32)
    assert np.allclose(vt2.apply_py(x), x @ A.T, atol=1e-6)
    with pytest.raises(RuntimeError):
        faiss.write_VectorTransform(faiss.OPQMatrix(12, 4, 8), fn)                 # untrained
