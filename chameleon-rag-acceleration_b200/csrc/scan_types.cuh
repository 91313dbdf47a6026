// scan_types.cuh -- the plain types and helpers shared by every scan kernel and by both translation units
// (api.cu, qlut.cu): kernel parameter blocks, pair-setup statistics, the oracle's non-fused squared difference.
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

#include "topk.cuh"

namespace b200 {

constexpr int kThreads = 256;

__device__ __forceinline__ float sqdiff_acc(float acc, float a, float b) {
    float diff = __fsub_rn(a, b);
    return __fadd_rn(acc, __fmul_rn(diff, diff));
}

struct PairStats {
    unsigned long long scan_codes;   // sum over valid pairs of list_size
    int nvalid;                      // number of pairs with a non-empty list
    int work_counter;                // dynamic scheduler of scan_pairs_kernel
    int ngroups;                     // number of same-list pair groups (scan_duo.cuh / scan_quad.cuh), sum over lists
                                     // of ceil(cnt / group size)
    int work_counter2;               // scheduler of a second pass over the same work items (scan_stream.cuh fallback)
    int pad_;
};

struct ScanParams {
    const float* xq;          // (nq, d)
    const float* cent;        // (nlist, d)
    const float* pq;          // (M, 256, dsub)
    const int64_t* offsets;   // (nlist + 1)
    const uint8_t* codes;     // (ntotal, M)
    const int32_t* probe;     // (nq * nprobe) list id or -1
    const int32_t* order;     // (nvalid) pair indices sorted by list
    const void* groups;       // (ngroups) work items: DuoGroup (scan_duo.cuh, scan_duo32.cuh) or QuadGroup (scan_quad.cuh)
    float4* lutf_scratch;     // scan_quad.cuh: one fp32 LUT set (16 x 256 float4) per CTA, global memory
    const float* pq_maxnorm;  // scan_quad.cuh: (M) max_c ||pq[m][c]||, slightly rounded up
    uint64_t* out_keys;       // (nq * nprobe, k)
    int* out_cnt;             // (nq * nprobe), pre-zeroed
    uint32_t* qthr;           // (nq) per-query threshold bits, pre-set to +inf
    PairStats* stats;
    int d, M, dsub, nprobe, k;
    const float* lutg;        // small batches: LUTs built once per (query, probe) pair by lut_small_kernel, or nullptr.
                              // Layout per pair: [c][m] (M == 16, the skewed kernel's store order) or [m][c]
    int quad_drain_at;        // scan_quad.cuh: survivors queued before the exact phase runs (<= 256)
    uint64_t negzero2;        // (-0.0f, -0.0f): an addend ptxas cannot see through (scan_duo.cuh lut_entry_duo)
    int nseg;                 // each (query, probe) pair is scanned by nseg CTAs (contiguous segments of its list):
                              // fills the GPU at small batch sizes; slot = pair * nseg + segment
};

}  // namespace b200
