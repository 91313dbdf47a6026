// coarse_small.cuh -- K1 for small batches (batch-1 latency path).
//
// For a handful of queries the tensor-core pre-filter (split, two GEMM passes, radix select, rescoring) is six
// launches of mostly latency.  Here the exact contract is computed directly in two launches:
//   coarse_small_dist_kernel    one CTA per 128 centroids: the rows are staged through shared memory in 32-dimension
//                               slices (coalesced float4 loads), every thread walks ITS centroid in the oracle's order
//                               (sum_j (q_j - c_j)^2, sequential, separately rounded) for up to 8 queries at once, and
//                               the CTA's 32 smallest (distance, id) keys per query are selected with the register
//                               sorting networks of topk.cuh;
//   coarse_small_select_kernel  one CTA per query: the nlist / 128 x 32 surviving keys -> the nprobe smallest, sorted.
// Exact for nprobe <= 32 (a CTA can contribute at most its 32 best).  Ties -> lower centroid id (unique 64-bit keys).
// Reference: IVFPQ_1B_search.ipynb:7922-7927, 7991-7999 (distance to every centroid, sort, take nprobe).
#pragma once
#include "kernels.cuh"

namespace b200 {

constexpr int kCsThreads = 128;
constexpr int kCsQ = 8;            // queries per pass
constexpr int kCsJ = 32;           // dimensions staged per step
constexpr int kCsKeep = 32;        // keys every CTA keeps per query
constexpr int kCsMaxNq = 16;
constexpr int64_t kCsMaxKeys = 8192;

inline bool coarse_small_usable(int64_t nq, int64_t nlist, int nprobe) {
    const int64_t nctas = (nlist + kCsThreads - 1) / kCsThreads;
    return nq <= kCsMaxNq && nprobe <= kCsKeep && nctas * kCsKeep <= kCsMaxKeys && nlist >= kCsThreads;
}

__host__ __device__ inline size_t coarse_small_dist_smem(int d) {
    return sizeof(float) * (static_cast<size_t>(kCsThreads) * (kCsJ + 1) + static_cast<size_t>(kCsQ) * ((d + 3) & ~3)) +
           sizeof(uint64_t) * kCsThreads;
}

__global__ void __launch_bounds__(kCsThreads)
coarse_small_dist_kernel(const float* __restrict__ xq, const float* __restrict__ cent, int nq, int64_t nlist, int d,
                         uint64_t* __restrict__ cand) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* tile = reinterpret_cast<float*>(smem_raw);                    // [128][33]
    const int dpad = (d + 3) & ~3;
    float* sq = tile + kCsThreads * (kCsJ + 1);                          // [kCsQ][dpad]
    uint64_t* runs = reinterpret_cast<uint64_t*>(sq + kCsQ * dpad);      // [4 warps][32]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t c0 = static_cast<int64_t>(blockIdx.x) * kCsThreads;
    const int q0 = blockIdx.y * kCsQ;
    const int nqc = min(kCsQ, nq - q0);
    const int64_t nctas = gridDim.x;
    for (int e = tid; e < nqc * d; e += kCsThreads) sq[(e / d) * dpad + e % d] = xq[static_cast<int64_t>(q0) * d + e];
    float acc[kCsQ];
#pragma unroll
    for (int q = 0; q < kCsQ; q++) acc[q] = 0.0f;
    const bool vec4 = (d & 3) == 0;
    const int64_t my_c = c0 + tid;
    const float* row = tile + tid * (kCsJ + 1);
    if (vec4) {
        // 8 threads x float4 cover one 128-byte row slice, 16 rows per pass; the loads of slice s + 2 are issued before
        // slice s is consumed (two register buffers), so the cold-memory latency overlaps the arithmetic
        const int c4 = (tid & 7) * 4, r0 = tid >> 3;
        const int nsl = (d + kCsJ - 1) / kCsJ;
        float4 va[kCsThreads / 16], vb[kCsThreads / 16];
#define CS_LOAD(V, S)                                                                                     \
    _Pragma("unroll") for (int ps = 0; ps < kCsThreads / 16; ps++) {                                      \
        const int64_t c = c0 + r0 + 16 * ps;                                                              \
        V[ps] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);                                                      \
        if ((S) < nsl && c < nlist && (S) * kCsJ + c4 < d)                                                \
            V[ps] = __ldg(reinterpret_cast<const float4*>(cent + c * d + (S) * kCsJ + c4));               \
    }
#define CS_CONSUME(V, S)                                                                                  \
    {                                                                                                     \
        __syncthreads(); /* previous slice consumed (first time round: the queries are staged) */         \
        _Pragma("unroll") for (int ps = 0; ps < kCsThreads / 16; ps++) {                                  \
            float* t = tile + (r0 + 16 * ps) * (kCsJ + 1) + c4;                                           \
            t[0] = V[ps].x;                                                                               \
            t[1] = V[ps].y;                                                                               \
            t[2] = V[ps].z;                                                                               \
            t[3] = V[ps].w;                                                                               \
        }                                                                                                 \
        __syncthreads();                                                                                  \
        CS_LOAD(V, (S) + 2)                                                                               \
        const int j0 = (S) * kCsJ, jn = min(kCsJ, d - j0);                                                \
        for (int jj = 0; jj < jn; jj++) {                                                                 \
            const float x = row[jj];                                                                      \
            _Pragma("unroll") for (int q = 0; q < kCsQ; q++)                                              \
                if (q < nqc) acc[q] = sqdiff_acc(acc[q], sq[q * dpad + j0 + jj], x);                      \
        }                                                                                                 \
    }
        CS_LOAD(va, 0)
        CS_LOAD(vb, 1)
        for (int s = 0; s < nsl; s += 2) {
            CS_CONSUME(va, s)
            if (s + 1 < nsl) CS_CONSUME(vb, s + 1)
        }
#undef CS_LOAD
#undef CS_CONSUME
    } else {
        for (int j0 = 0; j0 < d; j0 += kCsJ) {
            const int jn = min(kCsJ, d - j0);
            __syncthreads();
            for (int i = warp; i < kCsThreads; i += kCsThreads / 32) {
                const int64_t c = c0 + i;
                if (lane < jn) tile[i * (kCsJ + 1) + lane] = c < nlist ? __ldg(cent + c * d + j0 + lane) : 0.0f;
            }
            __syncthreads();
            for (int jj = 0; jj < jn; jj++) {
                const float x = row[jj];
#pragma unroll
                for (int q = 0; q < kCsQ; q++)
                    if (q < nqc) acc[q] = sqdiff_acc(acc[q], sq[q * dpad + j0 + jj], x);
            }
        }
    }
    // the CTA's 32 smallest keys per query: sort each warp's 32, merge the four runs
#pragma unroll 1
    for (int q = 0; q < nqc; q++) {
        float a = 0.0f;
#pragma unroll
        for (int qq = 0; qq < kCsQ; qq++)
            if (qq == q) a = acc[qq];
        uint64_t key = my_c < nlist ? make_key(__float_as_uint(a), static_cast<uint32_t>(my_c)) : kPadKey;
        key = TopK::warp_sort32(key, lane);
        __syncthreads();   // runs[] free again
        runs[warp * 32 + lane] = key;
        __syncthreads();
        if (warp == 0) {
            uint64_t v[1] = {runs[lane]};
            for (int w = 1; w < kCsThreads / 32; w++) {
                const uint64_t o[1] = {runs[w * 32 + lane]};
                TopK::warp_lower<1>(v, o, lane);
            }
            cand[(static_cast<int64_t>(q0 + q) * nctas + blockIdx.x) * kCsKeep + lane] = v[0];
        }
    }
}

// one CTA per query: n keys (n = nctas * 32 <= 8192) -> the nprobe smallest, ascending
__global__ void __launch_bounds__(kThreads)
coarse_small_select_kernel(const uint64_t* __restrict__ cand, int n, int nprobe, int32_t* __restrict__ probe32,
                           int64_t* __restrict__ ids64, float* __restrict__ dis_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* queue = reinterpret_cast<uint64_t*>(smem_raw);
    const int tid = threadIdx.x;
    const int64_t q = blockIdx.x;
    const int np = n > kThreads ? n : kThreads;   // reduce_queue parks the warps' runs in the first 256 slots
    for (int i = tid; i < np; i += kThreads) queue[i] = i < n ? cand[q * n + i] : kPadKey;
    __syncthreads();
    if (n > 32) {
        TopK::reduce_queue<1, kThreads>(queue, n);
    } else {
        if (tid < 32) queue[tid] = TopK::warp_sort32(queue[tid], tid);
        __syncthreads();
    }
    for (int i = tid; i < nprobe; i += kThreads) {
        const uint64_t key = queue[i];
        int32_t id = -1;
        float dv = FLT_MAX;
        if (key != kPadKey && static_cast<uint32_t>(key >> 32) < kInfBits) {
            id = static_cast<int32_t>(key & 0xffffffffu);
            dv = __uint_as_float(static_cast<uint32_t>(key >> 32));
        }
        if (probe32) probe32[q * nprobe + i] = id;
        if (ids64) ids64[q * nprobe + i] = id;
        if (dis_out) dis_out[q * nprobe + i] = dv;
    }
}

}  // namespace b200
