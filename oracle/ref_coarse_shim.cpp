// oracle/ref_coarse_shim.cpp -- TEST INFRASTRUCTURE ONLY (see ivfpq_oracle.c header).
//
// Runs the REFERENCE'S OWN coarse-quantizer code for row a1 of the hot path: the FPGA deployment's host selects the
// nprobe cells on the CPU with the vendored header-only hnswlib, exactly as
//   retrieval_accelerator/entire_accelerator_final_SIFT_M32/src/host.cpp:516-533  (BruteforceSearch over the nlist
//   centroids, L2Space(D), addPoint(centroid, id)) and :574-581 (searchKnn(query, nprobe), popped farthest-first).
// Nothing of the reference is copied: this file is ours, it only #includes the headers where they lie under
// /root/reference (oracle/Makefile passes -I), and the built library goes to oracle/_ref/ (git-ignored, travels to the
// GPU box).  It exists to PIN the oracle's coarse stage against code of the reference that runs here; the rest of the
// path (Faiss) stays unpinned.
//
// hnswlib's L2Sqr is SIMD (16 / 4 lanes, different summation order than the oracle's sequential contract), so the
// distances agree to rounding, not bit for bit, and exact ties may be ordered differently.
#include <cstdint>
#include <utility>
#include <vector>

#include "hnswlib/hnswlib.h"

extern "C" __attribute__((visibility("default")))
int ref_coarse_bruteforce(int64_t nlist, int d, const float* centroids, int64_t nq, const float* xq, int nprobe,
                          int64_t* ids, float* dis) {
    if (nprobe < 1 || nprobe > nlist) return -1;
    hnswlib::L2Space space(static_cast<size_t>(d));
    hnswlib::BruteforceSearch<float> alg(&space, static_cast<size_t>(nlist));
    for (int64_t i = 0; i < nlist; i++) alg.addPoint(centroids + static_cast<size_t>(i) * d, static_cast<size_t>(i));
    for (int64_t q = 0; q < nq; q++) {
        auto gd = alg.searchKnn(xq + static_cast<size_t>(q) * d, static_cast<size_t>(nprobe));
        if (static_cast<int>(gd.size()) != nprobe) return -2;
        int pos = nprobe;                     // the queue pops the farthest first: fill the row back to front
        while (!gd.empty()) {
            --pos;
            dis[q * nprobe + pos] = gd.top().first;
            ids[q * nprobe + pos] = static_cast<int64_t>(gd.top().second);
            gd.pop();
        }
    }
    return 0;
}

extern "C" __attribute__((visibility("default")))
const char* ref_coarse_simd(void) {
#if defined(USE_AVX512)
    return "avx512";
#elif defined(USE_AVX)
    return "avx";
#elif defined(USE_SSE)
    return "sse";
#else
    return "scalar";
#endif
}
