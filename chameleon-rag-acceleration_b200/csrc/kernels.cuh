// kernels.cuh -- sm_100a kernels of the IVF-PQ search path.
//
//   K1  coarse_dist_kernel (exact fp32 query x centroid L2^2; tensor-core path in coarse_tc.cuh, select in
//       select_radix.cuh)                           a1
//   K2+K3+K4  scan_pairs_kernel                      a2-a5 residual, LUT in shared memory, ADC scan, top-k
//   K4b merge_query_kernel                           a5,a6 per-query merge over probes + id lookup
//   K5  merge_shards_kernel                          multi-GPU merge after the all-gather
//   encode_kernel                                    a9  PQ encode of residuals for index.add
//
// Arithmetic follows the oracle contract bit for bit: every multiply/add is a separately rounded fp32
// op (__fsub_rn/__fmul_rn/__fadd_rn are never contracted into FMA) and every accumulation runs in the
// oracle's order (j ascending, m ascending).
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

#include "scan_types.cuh"

namespace b200 {

// ------------------------------------------------------------------------------------------------
// K1a: exact coarse distances.  out[q][c] = sum_j (xq[q][j] - cent[c][j])^2, j ascending.
// 64 x 64 tile per CTA, 4 x 4 per thread, operands staged through shared memory in k-chunks.
// Reference: IVFPQ_1B_search.ipynb:7922-7927, 7991-7996 (distance to every centroid).
// ------------------------------------------------------------------------------------------------
constexpr int kCoarseTile = 64;
constexpr int kCoarseKC = 16;

__global__ void __launch_bounds__(kThreads) coarse_dist_kernel(const float* __restrict__ xq,
                                                               const float* __restrict__ cent,
                                                               float* __restrict__ out, int nq, int64_t nlist,
                                                               int d, int64_t out_stride) {
    __shared__ __align__(16) float sQ[kCoarseKC][kCoarseTile + 4];
    __shared__ __align__(16) float sC[kCoarseKC][kCoarseTile + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t cbase = static_cast<int64_t>(blockIdx.x) * kCoarseTile;
    const int qbase = blockIdx.y * kCoarseTile;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0.0f;

    for (int k0 = 0; k0 < d; k0 += kCoarseKC) {
        const int kc = min(kCoarseKC, d - k0);
#pragma unroll
        for (int i = 0; i < (kCoarseTile * kCoarseKC) / kThreads; i++) {
            int e = tid + i * kThreads;
            int row = e / kCoarseKC, col = e % kCoarseKC;
            float qv = 0.0f, cv = 0.0f;
            if (col < kc) {
                if (qbase + row < nq) qv = xq[static_cast<int64_t>(qbase + row) * d + k0 + col];
                if (cbase + row < nlist) cv = cent[(cbase + row) * d + k0 + col];
            }
            sQ[col][row] = qv;
            sC[col][row] = cv;
        }
        __syncthreads();
        for (int kk = 0; kk < kc; kk++) {
            float4 qa = *reinterpret_cast<const float4*>(&sQ[kk][ty * 4]);
            float4 cb = *reinterpret_cast<const float4*>(&sC[kk][tx * 4]);
            float qv[4] = {qa.x, qa.y, qa.z, qa.w};
            float cv[4] = {cb.x, cb.y, cb.z, cb.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = sqdiff_acc(acc[i][j], qv[i], cv[j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int q = qbase + ty * 4 + i;
        if (q >= nq) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int64_t c = cbase + tx * 4 + j;
            if (c < nlist) out[static_cast<int64_t>(q) * out_stride + c] = acc[i][j];
        }
    }
}

// K1b (per-query nprobe-select) lives in select_radix.cuh.  Tile / queue sizes of the exact fallback kernel:
constexpr int kSelCap = 2048;
constexpr int kSelTile = kThreads * 4;

// convert caller-provided int64 list ids (search_preassigned) to the internal int32 probe table
__global__ void probes_from_i64_kernel(const int64_t* __restrict__ in, int32_t* __restrict__ out, int64_t n,
                                       int64_t nlist) {
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) {
        int64_t v = in[i];
        out[i] = (v >= 0 && v < nlist) ? static_cast<int32_t>(v) : -1;
    }
}

// ------------------------------------------------------------------------------------------------
// pair setup: order the (query, probe) pairs by list id so that CTAs scanning the same list run
// back to back and its codes are served from L2 after the first touch.  Counting sort in three tiny
// kernels; also accumulates the algorithmic scan statistics the roofline uses.
// ------------------------------------------------------------------------------------------------
__global__ void pair_hist_kernel(const int32_t* __restrict__ probe, int64_t npairs,
                                 const int64_t* __restrict__ offsets, int* __restrict__ hist,
                                 PairStats* __restrict__ stats) {
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    unsigned long long codes = 0;
    if (i < npairs) {
        int l = probe[i];
        if (l >= 0) {
            int64_t sz = offsets[l + 1] - offsets[l];
            if (sz > 0) {
                atomicAdd(&hist[l], 1);
                codes = static_cast<unsigned long long>(sz);
            }
        }
    }
    // warp-reduce then one atomic per warp
    for (int o = 16; o > 0; o >>= 1) codes += __shfl_down_sync(0xffffffffu, codes, o);
    if ((threadIdx.x & 31) == 0 && codes) atomicAdd(&stats->scan_codes, codes);
}

// exclusive scans of hist[nlist] -> start[nlist] (pairs) and of ceil(hist / 2) -> gstart[nlist] (two-query groups);
// hist is zeroed on the way out so that pair_scatter_kernel can use it as its per-list cursor.
// Single CTA (nlist <= a few 100k).
__global__ void __launch_bounds__(1024) pair_scan_kernel(int* __restrict__ hist, int* __restrict__ start,
                                                         int* __restrict__ gstart, int64_t nlist, int gsz,
                                                         PairStats* __restrict__ stats) {
    __shared__ int warp_sums[32], warp_gsums[32];
    __shared__ int carry, gcarry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
        carry = 0;
        gcarry = 0;
    }
    __syncthreads();
    for (int64_t base = 0; base < nlist; base += 1024) {
        int64_t i = base + tid;
        int v = i < nlist ? hist[i] : 0;
        int gv = (v + gsz - 1) / gsz;
        int x = v, gx = gv;
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, x, o);
            int gy = __shfl_up_sync(0xffffffffu, gx, o);
            if (lane >= o) {
                x += y;
                gx += gy;
            }
        }
        if (lane == 31) {
            warp_sums[wid] = x;
            warp_gsums[wid] = gx;
        }
        __syncthreads();
        if (wid == 0) {
            int w = warp_sums[lane], gw = warp_gsums[lane];
            for (int o = 1; o < 32; o <<= 1) {
                int y = __shfl_up_sync(0xffffffffu, w, o);
                int gy = __shfl_up_sync(0xffffffffu, gw, o);
                if (lane >= o) {
                    w += y;
                    gw += gy;
                }
            }
            warp_sums[lane] = w;
            warp_gsums[lane] = gw;
        }
        __syncthreads();
        int prefix = carry + (wid ? warp_sums[wid - 1] : 0) + x - v;
        int gprefix = gcarry + (wid ? warp_gsums[wid - 1] : 0) + gx - gv;
        if (i < nlist) {
            start[i] = prefix;
            gstart[i] = gprefix;
            hist[i] = 0;
        }
        __syncthreads();
        if (tid == 1023) {
            carry = prefix + v;
            gcarry = gprefix + gv;
        }
        __syncthreads();
    }
    if (tid == 0) {
        stats->nvalid = carry;
        stats->ngroups = gcarry;
        stats->work_counter = 0;
        stats->work_counter2 = 0;
    }
}

// One work item of the two-query scan (scan_duo.cuh): two (query, probe) pairs of the same list plus everything the
// kernel needs to start on it, so that it is one 32-byte load away (no groups -> probe -> offsets pointer chase).
struct __align__(16) DuoGroup {
    int pair_a;
    int pair_b;     // -1: the list had an odd number of probing queries
    int list;
    uint32_t n;     // list length (> 0)
    int64_t beg;    // first row of the list in codes / ids
    int64_t pad_;
};
static_assert(sizeof(DuoGroup) == 32, "DuoGroup is two 16-byte words");

// One work item of the four-query filter scan (scan_quad.cuh): up to four pairs of the same list.
struct __align__(16) QuadGroup {
    int pair[4];    // -1: unused slot (always at the end)
    int list;
    uint32_t n;     // list length (> 0)
    int64_t beg;    // first row of the list in codes / ids
    int64_t pad_[2];
};
static_assert(sizeof(QuadGroup) == 48, "QuadGroup is three 16-byte words");

// order[] = pair indices sorted by list; groups[] = the same pairs two by two (pair_b is pre-set to -1 by a 0xff
// memset, so the odd pair of a list keeps it).  Which queries end up together depends on the atomics' order; the
// results do not.
__global__ void pair_scatter_kernel(const int32_t* __restrict__ probe, int64_t npairs,
                                    const int64_t* __restrict__ offsets, const int* __restrict__ start,
                                    const int* __restrict__ gstart, int* __restrict__ cursor,
                                    int32_t* __restrict__ order, void* __restrict__ groups, int gsz) {
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= npairs) return;
    int l = probe[i];
    if (l < 0) return;
    const int64_t beg = offsets[l], sz = offsets[l + 1] - beg;
    if (sz <= 0) return;
    const int rank = atomicAdd(&cursor[l], 1);
    order[start[l] + rank] = static_cast<int32_t>(i);
    if (gsz == 4) {
        QuadGroup* g = static_cast<QuadGroup*>(groups) + gstart[l] + (rank >> 2);
        g->pair[rank & 3] = static_cast<int32_t>(i);
        if ((rank & 3) == 0) {
            g->list = l;
            g->n = static_cast<uint32_t>(sz);
            g->beg = beg;
        }
        return;
    }
    DuoGroup* g = static_cast<DuoGroup*>(groups) + gstart[l] + (rank >> 1);
    if (rank & 1) {
        g->pair_b = static_cast<int32_t>(i);
    } else {
        g->pair_a = static_cast<int32_t>(i);
        g->list = l;
        g->n = static_cast<uint32_t>(sz);
        g->beg = beg;
    }
}

// Small batches (a few thousand pairs at most): nothing is gained by sorting the pairs by list, so the whole pair setup
// is ONE single-CTA kernel: valid pairs are compacted in their natural order (stable), the per-query thresholds are
// set to +inf, the per-slot candidate counts to zero, and the scan statistics are accumulated.
__global__ void __launch_bounds__(1024) pair_setup_small_kernel(const int32_t* __restrict__ probe, int npairs,
                                                                const int64_t* __restrict__ offsets,
                                                                int32_t* __restrict__ order, uint32_t* __restrict__ qthr,
                                                                int nq, int* __restrict__ out_cnt, int nslots,
                                                                PairStats* __restrict__ stats) {
    __shared__ int warp_sums[32];
    __shared__ int carry;
    __shared__ unsigned long long total_codes;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
        carry = 0;
        total_codes = 0ull;
    }
    for (int i = tid; i < nq; i += 1024) qthr[i] = kInfBits;
    for (int i = tid; i < nslots; i += 1024) out_cnt[i] = 0;
    __syncthreads();
    for (int base = 0; base < npairs; base += 1024) {
        const int i = base + tid;
        unsigned long long codes = 0ull;
        if (i < npairs) {
            const int l = probe[i];
            if (l >= 0) {
                const int64_t sz = offsets[l + 1] - offsets[l];
                if (sz > 0) codes = static_cast<unsigned long long>(sz);
            }
        }
        const bool valid = codes != 0ull;
        const unsigned bal = __ballot_sync(0xffffffffu, valid);
        if (lane == 0) warp_sums[wid] = __popc(bal);
        unsigned long long csum = codes;
        for (int o = 16; o > 0; o >>= 1) csum += __shfl_down_sync(0xffffffffu, csum, o);
        if (lane == 0 && csum) atomicAdd(&total_codes, csum);
        __syncthreads();
        if (wid == 0) {
            int w = warp_sums[lane];
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            warp_sums[lane] = w;   // inclusive
        }
        __syncthreads();
        const int pos = carry + (wid ? warp_sums[wid - 1] : 0) + __popc(bal & lanemask_lt());
        if (valid) order[pos] = i;
        __syncthreads();
        if (tid == 0) carry += warp_sums[31];
        __syncthreads();
    }
    if (tid == 0) {
        stats->nvalid = carry;
        stats->ngroups = 0;
        stats->work_counter = 0;
        atomicAdd(&stats->scan_codes, total_codes);
    }
}

__global__ void fill_u32_kernel(uint32_t* __restrict__ p, int64_t n, uint32_t v) {
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ------------------------------------------------------------------------------------------------
// K2+K3+K4: one CTA per (query, probe) pair, persistent CTAs pulling pairs from a work counter.
//   a2 residual  r = q - c                                   (ipynb:8006, LUT_construction.hpp:182-187)
//   a3 LUT       T[m][k] = sum_j (r[m*dsub+j] - pq[m][k][j])^2 in shared memory
//                                                            (ipynb:7929-7946, LUT_construction.hpp:189-209)
//   a4 ADC       dist = sum_m T[m][code[m]], m ascending       (ipynb:7948-7960, ADC.hpp:88-91)
//   a5 top-k     exact, (distance, offset) order, per pair    (ipynb:7980-7982)
// Generic in M (any M with d % M == 0); codes are read straight from HBM/L2 into registers with the
// widest aligned vector load M allows (VEC bytes).  A per-query global threshold (the best known k-th
// distance over the probes finished so far) prunes candidates across the probes of a query.
// ------------------------------------------------------------------------------------------------
// Small batches split every (query, probe) pair into nseg list segments, one CTA each; building the pair's LUT in each
// of those CTAs multiplies the LUT work and the PQ-codebook traffic by nseg (C4 shape: 786 KB of codebook per CTA).
// This kernel builds every pair's LUT ONCE: grid (pair, m-chunk of 8 sub-quantizers), thread = code value.
//   T[m][c] = sum_j ((q - cent)[m*dsub+j] - pq[m][c][j])^2, j ascending, separately rounded (the oracle's form).
__global__ void __launch_bounds__(256) lut_small_kernel(const float* __restrict__ xq, const float* __restrict__ cent,
                                                        const float* __restrict__ pq, const int32_t* __restrict__ probe,
                                                        int nprobe, int d, int M, int dsub, int cm_layout,
                                                        float* __restrict__ lutg) {
    const int pair = blockIdx.x, c = threadIdx.x;
    const int list = probe[pair];
    if (list < 0) return;
    const int q = pair / nprobe;
    const float* xr = xq + static_cast<int64_t>(q) * d;
    const float* cr = cent + static_cast<int64_t>(list) * d;
    float* out = lutg + static_cast<int64_t>(pair) * M * 256;
    const int m0 = blockIdx.y * 8, m1 = min(M, m0 + 8);
    for (int m = m0; m < m1; m++) {
        const float* pc = pq + (static_cast<int64_t>(m) * 256 + c) * dsub;
        float acc = 0.0f;
        for (int j = 0; j < dsub; j++) {
            const float r = __fsub_rn(__ldg(xr + m * dsub + j), __ldg(cr + m * dsub + j));
            acc = sqdiff_acc(acc, r, __ldg(pc + j));
        }
        out[cm_layout ? c * M + m : m * 256 + c] = acc;
    }
}

constexpr int kScanCap = 2048;
constexpr int kScanUnroll = 4;
constexpr int kScanTile = kThreads * kScanUnroll;

template <int VEC>
__device__ __forceinline__ float adc_one(const float* __restrict__ lut, const uint8_t* __restrict__ code, int M);

template <>
__device__ __forceinline__ float adc_one<16>(const float* __restrict__ lut, const uint8_t* __restrict__ code,
                                             int M) {
    float acc = 0.0f;
    const uint4* p = reinterpret_cast<const uint4*>(code);
    for (int m0 = 0; m0 < M; m0 += 16) {
        uint4 v = __ldg(p + (m0 >> 4));
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int b = 0; b < 4; b++)
                acc = __fadd_rn(acc, lut[(m0 + i * 4 + b) * 256 + ((w[i] >> (8 * b)) & 0xffu)]);
    }
    return acc;
}

template <>
__device__ __forceinline__ float adc_one<8>(const float* __restrict__ lut, const uint8_t* __restrict__ code,
                                            int M) {
    float acc = 0.0f;
    const uint2* p = reinterpret_cast<const uint2*>(code);
    for (int m0 = 0; m0 < M; m0 += 8) {
        uint2 v = __ldg(p + (m0 >> 3));
        uint32_t w[2] = {v.x, v.y};
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int b = 0; b < 4; b++)
                acc = __fadd_rn(acc, lut[(m0 + i * 4 + b) * 256 + ((w[i] >> (8 * b)) & 0xffu)]);
    }
    return acc;
}

template <>
__device__ __forceinline__ float adc_one<4>(const float* __restrict__ lut, const uint8_t* __restrict__ code,
                                            int M) {
    float acc = 0.0f;
    const uint32_t* p = reinterpret_cast<const uint32_t*>(code);
    for (int m0 = 0; m0 < M; m0 += 4) {
        uint32_t w = __ldg(p + (m0 >> 2));
#pragma unroll
        for (int b = 0; b < 4; b++) acc = __fadd_rn(acc, lut[(m0 + b) * 256 + ((w >> (8 * b)) & 0xffu)]);
    }
    return acc;
}

template <>
__device__ __forceinline__ float adc_one<1>(const float* __restrict__ lut, const uint8_t* __restrict__ code,
                                            int M) {
    float acc = 0.0f;
    for (int m = 0; m < M; m++) acc = __fadd_rn(acc, lut[m * 256 + __ldg(code + m)]);
    return acc;
}

// shared memory: [ LUT M*256 f32 | residual d f32 | TopK ]
__host__ __device__ inline size_t scan_smem_bytes(int M, int d, int k) {
    size_t lut = sizeof(float) * static_cast<size_t>(M) * 256;
    size_t res = sizeof(float) * static_cast<size_t>((d + 3) & ~3);
    return lut + res + TopK::smem_bytes(k, kScanCap);
}

template <int VEC>
__global__ void __launch_bounds__(kThreads) scan_pairs_kernel(const ScanParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* lut = reinterpret_cast<float*>(smem_raw);
    float* res = lut + p.M * 256;
    TopK tk;
    tk.bind(res + ((p.d + 3) & ~3), p.k, kScanCap);
    __shared__ int s_work;
    const int tid = threadIdx.x;
    const int nvalid = p.stats->nvalid;

    int next_work = 0;
    if (tid == 0) next_work = atomicAdd(&p.stats->work_counter, 1);
    for (;;) {
        if (tid == 0) s_work = next_work;
        __syncthreads();
        const int w = s_work;
        if (w >= nvalid * p.nseg) break;
        if (tid == 0) next_work = atomicAdd(&p.stats->work_counter, 1);   // prefetch the next work item
        const int pair = p.order[w / p.nseg];
        const int seg = w % p.nseg;
        const int slot = pair * p.nseg + seg;
        const int q = pair / p.nprobe;
        const int list = p.probe[pair];
        const int64_t beg = p.offsets[list];
        const int64_t ntot = p.offsets[list + 1] - beg;
        const int64_t seglen = (((ntot + p.nseg - 1) / p.nseg) + 255) & ~static_cast<int64_t>(255);
        const int64_t soff = seg * seglen;                        // first code of this segment
        if (soff >= ntot) {                                       // empty segment: out_cnt stays 0
            __syncthreads();
            continue;
        }
        const int64_t n = min(seglen, ntot - soff);

        const uint32_t ext_thr = *reinterpret_cast<volatile uint32_t*>(p.qthr + q);
        if (tid == 0) tk.reset(ext_thr);
        if (p.lutg) {
            // small batches: the pair's LUT was built once by lut_small_kernel ([m][c] layout)
            const float4* src = reinterpret_cast<const float4*>(p.lutg + static_cast<int64_t>(pair) * p.M * 256);
            float4* dst = reinterpret_cast<float4*>(lut);
            for (int idx = tid; idx < p.M * 64; idx += kThreads) dst[idx] = src[idx];
        } else {
            // a2: residual
            for (int j = tid; j < p.d; j += kThreads)
                res[j] = __fsub_rn(p.xq[static_cast<int64_t>(q) * p.d + j], p.cent[static_cast<int64_t>(list) * p.d + j]);
            __syncthreads();
            // a3: LUT
            for (int idx = tid; idx < p.M * 256; idx += kThreads) {
                const int m = idx >> 8;
                const float* pc = p.pq + static_cast<int64_t>(idx) * p.dsub;
                const float* r = res + m * p.dsub;
                float acc = 0.0f;
                for (int j = 0; j < p.dsub; j++) acc = sqdiff_acc(acc, r[j], __ldg(pc + j));
                lut[idx] = acc;
            }
        }
        __syncthreads();
        // a4 + a5
        const uint8_t* lcodes = p.codes + (beg + soff) * p.M;
        uint32_t thr = ext_thr;
        for (int64_t base = 0; base < n; base += kScanTile) {
#pragma unroll
            for (int u = 0; u < kScanUnroll; u++) {
                int64_t i = base + u * kThreads + tid;
                uint32_t bits = 0xffffffffu;
                if (i < n) bits = __float_as_uint(adc_one<VEC>(lut, lcodes + i * p.M, p.M));
                tk.push(bits <= thr, make_key(bits, static_cast<uint32_t>(soff + i)));
            }
            tk.sync_and_flush_if_over<kThreads>(kScanCap - kScanTile, ext_thr);
            thr = tk.threshold();
        }
        __syncthreads();
        tk.flush<kThreads>(ext_thr);
        const int nb = tk.count();
        const uint64_t* s = tk.sorted();
        for (int i = tid; i < nb; i += kThreads) p.out_keys[static_cast<int64_t>(slot) * p.k + i] = s[i];
        if (tid == 0) {
            p.out_cnt[slot] = nb;
            if (nb == p.k) atomicMin(p.qthr + q, static_cast<uint32_t>(s[p.k - 1] >> 32));
        }
        __syncthreads();   // smem reuse by the next pair
    }
}

// ------------------------------------------------------------------------------------------------
// K4b: per-query merge of the per-probe candidate lists + id lookup (a5 final order, a6).
// Total order (distance, probe rank, offset): within a probe the list is already (distance, offset)
// sorted, so position j stands in for the offset.  ids: ipynb:756-771 (get_invlist).
// ------------------------------------------------------------------------------------------------
constexpr int kMergeCap = 2048;
constexpr int kMergeTile = kThreads * 4;

__global__ void __launch_bounds__(kThreads) merge_query_kernel(const uint64_t* __restrict__ pair_keys,
                                                               const int* __restrict__ pair_cnt,
                                                               const int32_t* __restrict__ probe,
                                                               const int64_t* __restrict__ offsets,
                                                               const int64_t* __restrict__ ids, int nprobe, int k,
                                                               int nseg, const uint32_t* __restrict__ qthr,
                                                               float* __restrict__ D, int64_t* __restrict__ I,
                                                               const int* __restrict__ guard,
                                                               const int* __restrict__ qflag) {
    // guard: a fallback launch (scan_stream.cuh): all queries if (*guard & 3), else only the queries with qflag[q] != 0
    if (guard) {
        const int ov = *reinterpret_cast<const volatile int*>(guard);
        if (ov == 0 || ((ov & 3) == 0 && qflag[blockIdx.x] == 0)) return;
    }
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TopK tk;
    tk.bind(smem_raw, k, kMergeCap);
    const int tid = threadIdx.x;
    const int64_t q = blockIdx.x;
    // qthr[q] = smallest k-th-best distance any (probe, segment) of this query reached: an upper bound on the
    // final k-th distance, so candidates above it (strictly) can be dropped before any sorting
    const uint32_t ext_thr = qthr ? qthr[q] : kInfBits;
    if (tid == 0) tk.reset(ext_thr);
    __syncthreads();
    uint32_t thr = ext_thr;
    // slots of a query: (probe rank, segment) in scan order, so slot * k + j still orders like (rank, offset)
    const int nslot = nprobe * nseg;
    const int64_t total = static_cast<int64_t>(nslot) * k;
    for (int64_t base = 0; base < total; base += kMergeTile) {
#pragma unroll
        for (int u = 0; u < kMergeTile / kThreads; u++) {
            int64_t c = base + u * kThreads + tid;
            uint32_t bits = 0xffffffffu;
            if (c < total) {
                int pr = static_cast<int>(c / k), j = static_cast<int>(c % k);
                int64_t slot = q * nslot + pr;
                if (j < pair_cnt[slot]) bits = static_cast<uint32_t>(pair_keys[slot * k + j] >> 32);
            }
            tk.push(bits <= thr, make_key(bits, static_cast<uint32_t>(c)));
        }
        tk.sync_and_flush_if_over<kThreads>(kMergeCap - kMergeTile, ext_thr);
        thr = tk.threshold();
    }
    __syncthreads();
    tk.flush<kThreads>(ext_thr);
    const int nb = tk.count();
    const uint64_t* s = tk.sorted();
    for (int i = tid; i < k; i += kThreads) {
        float dv = FLT_MAX;
        int64_t id = -1;
        if (i < nb) {
            uint32_t tag = static_cast<uint32_t>(s[i] & 0xffffffffu);
            int pr = tag / k, j = tag % k;
            int64_t slot = q * nslot + pr;
            uint32_t off = static_cast<uint32_t>(pair_keys[slot * k + j] & 0xffffffffu);
            int64_t pos = offsets[probe[q * nprobe + pr / nseg]] + off;
            id = ids ? ids[pos] : pos;
            dv = __uint_as_float(static_cast<uint32_t>(s[i] >> 32));
        }
        D[q * k + i] = dv;
        I[q * k + i] = id;
    }
}

// ------------------------------------------------------------------------------------------------
// K5: merge of per-shard results after the all-gather.  Ds/Is are (nshard, nq, k); order
// (distance, shard, position) = concatenate + stable argsort + take k
// (bench_multi_cpu_performance_OSDI.py:203-219).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) merge_shards_kernel(const float* __restrict__ Ds,
                                                                const int64_t* __restrict__ Is, int nshard,
                                                                int64_t nq, int k, float* __restrict__ D,
                                                                int64_t* __restrict__ I) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TopK tk;
    tk.bind(smem_raw, k, kMergeCap);
    const int tid = threadIdx.x;
    const int64_t q = blockIdx.x;
    if (tid == 0) tk.reset(kInfBits);
    __syncthreads();
    uint32_t thr = kInfBits;
    const int64_t total = static_cast<int64_t>(nshard) * k;
    for (int64_t base = 0; base < total; base += kMergeTile) {
#pragma unroll
        for (int u = 0; u < kMergeTile / kThreads; u++) {
            int64_t c = base + u * kThreads + tid;
            uint32_t bits = 0xffffffffu;
            if (c < total) {
                int64_t s = c / k, j = c % k;
                int64_t src = (s * nq + q) * k + j;
                if (Is[src] >= 0) bits = __float_as_uint(Ds[src]);
            }
            tk.push(bits <= thr, make_key(bits, static_cast<uint32_t>(c)));
        }
        tk.sync_and_flush_if_over<kThreads>(kMergeCap - kMergeTile, kInfBits);
        thr = tk.threshold();
    }
    __syncthreads();
    tk.flush<kThreads>(kInfBits);
    const int nb = tk.count();
    const uint64_t* s = tk.sorted();
    for (int i = tid; i < k; i += kThreads) {
        float dv = FLT_MAX;
        int64_t id = -1;
        if (i < nb) {
            uint32_t tag = static_cast<uint32_t>(s[i] & 0xffffffffu);
            int64_t sh = tag / k, j = tag % k;
            int64_t src = (sh * nq + q) * k + j;
            id = Is[src];
            dv = Ds[src];
        }
        D[q * k + i] = dv;
        I[q * k + i] = id;
    }
}

// K5 over peer memory: the same merge, but every shard's (D, I) is read in place from the GPU that produced it --
// bufs[s] is rank s's result buffer mapped into this GPU's address space (NVLink / NVSwitch peer access), D at byte
// offset d_off and I at i_off.  Replaces "NCCL all-gather into a staging tensor, unpack, merge" by one kernel whose
// loads ARE the exchange (nq * k * 12 bytes per peer).  The caller brackets it with the symmetric-memory barriers.
__global__ void __launch_bounds__(kThreads) merge_shards_peer_kernel(const unsigned char* const* __restrict__ bufs,
                                                                     int64_t d_off, int64_t i_off, int nshard,
                                                                     int64_t nq, int k, float* __restrict__ D,
                                                                     int64_t* __restrict__ I) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TopK tk;
    tk.bind(smem_raw, k, kMergeCap);
    const int tid = threadIdx.x;
    const int64_t q = blockIdx.x;
    if (tid == 0) tk.reset(kInfBits);
    __syncthreads();
    uint32_t thr = kInfBits;
    const int64_t total = static_cast<int64_t>(nshard) * k;
    for (int64_t base = 0; base < total; base += kMergeTile) {
#pragma unroll
        for (int u = 0; u < kMergeTile / kThreads; u++) {
            int64_t c = base + u * kThreads + tid;
            uint32_t bits = 0xffffffffu;
            if (c < total) {
                const int64_t s = c / k, j = c % k;
                const int64_t* Is = reinterpret_cast<const int64_t*>(bufs[s] + i_off);
                const float* Ds = reinterpret_cast<const float*>(bufs[s] + d_off);
                // both remote loads are issued together (one NVLink round trip, not two)
                const int64_t id = __ldcv(Is + q * k + j);
                const float dv = __ldcv(Ds + q * k + j);
                if (id >= 0) bits = __float_as_uint(dv);
            }
            tk.push(bits <= thr, make_key(bits, static_cast<uint32_t>(c)));
        }
        tk.sync_and_flush_if_over<kThreads>(kMergeCap - kMergeTile, kInfBits);
        thr = tk.threshold();
    }
    __syncthreads();
    tk.flush<kThreads>(kInfBits);
    const int nb = tk.count();
    const uint64_t* s = tk.sorted();
    for (int i = tid; i < k; i += kThreads) {
        float dv = FLT_MAX;
        int64_t id = -1;
        if (i < nb) {
            const uint32_t tag = static_cast<uint32_t>(s[i] & 0xffffffffu);
            const int64_t sh = tag / k, j = tag % k;
            id = __ldcv(reinterpret_cast<const int64_t*>(bufs[sh] + i_off) + q * k + j);
            dv = __ldcv(reinterpret_cast<const float*>(bufs[sh] + d_off) + q * k + j);
        }
        D[q * k + i] = dv;
        I[q * k + i] = id;
    }
}

// ------------------------------------------------------------------------------------------------
// a9: PQ encode of residuals for index.add.  grid = (ceil(n / 256), M); each CTA stages pq[m] in
// shared memory, each thread encodes sub-vector m of one vector: argmin_k sum_j ((x-c)_j - pq[m][k][j])^2,
// ties -> lower k.  Same arithmetic as the LUT (the code is the argmin of the vector's own LUT row).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) encode_kernel(const float* __restrict__ x,
                                                          const float* __restrict__ cent,
                                                          const int64_t* __restrict__ list_no,
                                                          const float* __restrict__ pq, int64_t n, int d, int M,
                                                          int dsub, uint8_t* __restrict__ codes) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* spq = reinterpret_cast<float*>(smem_raw);              // [256 * dsub]
    float* sres = spq + 256 * dsub;                               // [dsub][kThreads]
    const int tid = threadIdx.x;
    const int m = blockIdx.y;
    const int64_t i = static_cast<int64_t>(blockIdx.x) * kThreads + tid;
    for (int e = tid; e < 256 * dsub; e += kThreads) spq[e] = pq[static_cast<int64_t>(m) * 256 * dsub + e];
    if (i < n) {
        const float* xv = x + i * d + m * dsub;
        const float* cv = cent + list_no[i] * d + m * dsub;
        for (int j = 0; j < dsub; j++) sres[j * kThreads + tid] = __fsub_rn(xv[j], cv[j]);
    }
    __syncthreads();
    if (i >= n) return;
    float best = 0.0f;
    int arg = -1;
    for (int kk = 0; kk < 256; kk++) {
        float acc = 0.0f;
        for (int j = 0; j < dsub; j++) acc = sqdiff_acc(acc, sres[j * kThreads + tid], spq[kk * dsub + j]);
        if (arg < 0 || acc < best) {
            best = acc;
            arg = kk;
        }
    }
    codes[i * M + m] = static_cast<uint8_t>(arg);
}

// ------------------------------------------------------------------------------------------------
// a10 (train): the centroid update of Lloyd's iteration as a DETERMINISTIC segmented sum.  Rows are visited in
// the order given (a stable sort of the assignment), one CTA per centroid, one column per thread, sequential fp32
// adds: the same training set and seeds give the same codebooks bit for bit on every run (an atomic scatter-add does
// not, and the bench lines of different runs / GPU counts would then come from different indexes).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) segment_sums_kernel(const float* __restrict__ x,
                                                                const int64_t* __restrict__ order,
                                                                const int64_t* __restrict__ start, int d,
                                                                float* __restrict__ sums) {
    const int64_t c = blockIdx.x;
    const int64_t beg = start[c], end = start[c + 1];
    for (int j = threadIdx.x; j < d; j += kThreads) {
        float a = 0.0f;
        for (int64_t i = beg; i < end; i++) a = __fadd_rn(a, __ldg(x + order[i] * d + j));
        sums[c * d + j] = a;
    }
}

}  // namespace b200
