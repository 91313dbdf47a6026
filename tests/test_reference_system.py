"""The whole hot path against the REFERENCE'S OWN SYSTEM run here, end to end (rows a1-a6):

  host:   hnswlib BruteforceSearch picks the nprobe cells           (host.cpp:516-581; oracle/ref_coarse_shim.cpp)
  kernel: `vadd` -- packet parsing, LUT construction, ADC PEs over four DRAM banks, hierarchical priority queue with the
          vector-id lookup, result packing                          (vadd.cpp; oracle/ref_accel_shim.cpp, C simulation)

compiled from the sources under /root/reference into oracle/_ref/ (built here, travels prebuilt to the GPU box).  The DRAM
images are laid out by oracle.ref_accel_search from the formats the kernel parses.  TOPK = 100 is the reference's
compile-time constant.  The accelerator's top-k is approximate by design (truncated first-level queues); the cases here
scan ~2000 uniformly coded vectors per query, where no first-level queue can hold more than its 15 / 23 slots of the
best 100 -- and the comparison itself would show it if one did."""
import numpy as np
import pytest

import _util

VARIANTS = ["SIFT_M16", "SIFT_M32", "Deep_M16", "Deep_M32"]


@pytest.fixture(scope="module")
def ref(oracle):
    if oracle.build_ref() is None:
        pytest.skip("oracle/_ref not available (reference not mounted and nothing prebuilt)")
    try:
        for v in VARIANTS:
            oracle.ref_accel_dims(v)
    except FileNotFoundError as e:
        pytest.skip(str(e))
    return oracle


def _index(oracle, variant, seed, scale):
    D, M, topk, _, _ = oracle.ref_accel_dims(variant)
    rng = np.random.default_rng(seed)
    nlist, nq = 24, (6 if seed % 2 else 40)          # a latency-path batch (list segments) and a throughput-path one
    pq = (rng.standard_normal((M, 256, D // M)) * 0.3 * scale).astype(np.float32)
    cent = (rng.random((nlist, D), dtype=np.float32) * np.float32(scale)).astype(np.float32)
    xq = (rng.random((nq, D), dtype=np.float32) * np.float32(scale)).astype(np.float32)
    offsets = np.zeros(nlist + 1, np.int64)
    offsets[1:] = np.cumsum(rng.integers(180, 330, nlist))
    n = int(offsets[-1])
    codes = rng.integers(0, 256, (n, M), dtype=np.uint8)
    ids = (rng.permutation(n) + 10 ** 6).astype(np.int64)          # user ids, not positions
    return dict(D=D, M=M, k=topk, nprobe=8, pq=pq, cent=cent, xq=xq, offsets=offsets, codes=codes, ids=ids)


def _reference_system(oracle, variant, c):
    _, probes = oracle.ref_coarse(c["xq"], c["cent"], c["nprobe"])                      # the host's cell selection
    return oracle.ref_accel_search(variant, c["cent"], c["pq"], c["offsets"], c["codes"], c["ids"], c["xq"], probes)


def _check(oracle, variant, seed, scale, search):
    c = _index(oracle, variant, seed, scale)
    Dr, Ir = _reference_system(oracle, variant, c)
    D, I = search(c)
    assert len({float(x) for x in Dr.ravel()}) > Dr.size * 0.9, "the case is meant to be (nearly) free of ties"
    _util.assert_same_modulo_ties(np.asarray(D, np.float32), np.asarray(I), Dr, Ir, f"{variant} seed {seed}")


@pytest.mark.parametrize("seed,scale", [(5, 1.0), (6, 255.0)])
@pytest.mark.parametrize("variant", VARIANTS)
def test_oracle_search_equals_the_references_host_plus_accelerator(ref, variant, seed, scale):
    def run(c):
        return ref.C.search(c["xq"], c["cent"], c["pq"], c["offsets"], c["codes"], c["ids"], c["nprobe"], c["k"])
    _check(ref, variant, seed, scale, run)


def test_accelerator_needs_filled_first_level_queues(ref):
    c = _index(ref, "SIFT_M16", 5, 1.0)
    with pytest.raises(ValueError):
        ref.ref_accel_search("SIFT_M16", c["cent"], c["pq"], c["offsets"], c["codes"], c["ids"], c["xq"],
                             np.zeros((6, 1), np.int64))


@pytest.mark.gpu
@pytest.mark.parametrize("seed,scale", [(5, 1.0), (6, 255.0)])
@pytest.mark.parametrize("variant", VARIANTS)
def test_cuda_search_equals_the_references_host_plus_accelerator(ref, variant, seed, scale):
    """index.search on the B200 against the reference's own deployment path, distances bit for bit."""
    import b200ivfpq as faiss

    def run(c):
        index = faiss.IndexIVFPQ(faiss.IndexFlatL2(c["D"]), c["D"], c["cent"].shape[0], c["M"], 8)
        index.set_codebooks(c["cent"], c["pq"])
        index.set_lists(c["offsets"], c["codes"], c["ids"])
        index.nprobe = c["nprobe"]
        return index.search(c["xq"], c["k"])
    _check(ref, variant, seed, scale, run)
