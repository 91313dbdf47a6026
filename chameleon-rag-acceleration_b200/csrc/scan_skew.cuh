// scan_skew.cuh -- bank-conflict-free ADC scan for M = 16 (K2+K3+K4, the hot kernel of the path).
//
// Why.  The ADC inner loop is one shared-memory gather per code byte.  With the textbook mapping
// (lane = code, all lanes on sub-quantizer m at the same time) the 32 lanes of a warp hit banks
// code[m] % 32 -- random -- and every LDS replays ~3.5x.  At 10k-query batches the codes of a list are
// served from L2 to ~39 queries, so the kernel is bound by that shared-memory gather, not by HBM.
//
// How.  Lanes are skewed in TIME instead: at step p lane l looks up sub-quantizer m = (l + p) % 16 of
// the code it is currently on, from a LUT whose rows are indexed by code value and whose columns are
// the sub-quantizers laid out periodically:
//         lut[c][w] = T[w % 16][c],   w = 0..47,  row stride 64 words (256 B)
// Lane l reads word w = l + p of row c: bank (l + p) % 32 -- all 32 lanes distinct, for ANY codes.
// Each lane still adds its code's 16 table entries in ascending m (m wraps to 0 exactly when the lane
// moves on to its next code), so distances are bit-identical to the oracle's sequential sum.
//   - the lane's 16 lookups per block read bytes r..15 of its current code and bytes 0..r-1 of its
//     next one (r = l % 16): a byte-rotated 16-byte window built once per block from registers;
//   - "restart at m = 0" and "capture at m = 15" happen at lane-dependent steps; they are folded into
//     the adds as exact FMAs with lane-constant 0/1 multipliers:  acc = acc*keep_p + T,
//     fin = acc*cap_p + fin  (x*1 + y and x*0 + y round exactly like y + x and y);
//   - lookup address = (c << 8) | (4*l) in ONE PRMT; the step offset 4*p is an LDS immediate.
// Per lookup: PRMT, LDS, FFMA, FFMA.
//
// Reference semantics: ADC.hpp:75-99 / IVFPQ_1B_search.ipynb:7948-7960 (sum over m ascending),
// LUT_construction.hpp:180-209 / ipynb:7929-7946 (LUT), priority_queue_L1.hpp:65-75 (strict <).
#pragma once
#include "kernels.cuh"

namespace b200 {

constexpr int kSkewRowWords = 64;
constexpr int kSkewLutBytes = 256 * kSkewRowWords * 4;   // 64 KB
constexpr int kSkewTB = 4;                                // blocks (of 32 codes) per warp per tile
static_assert(kThreads * kSkewTB <= kScanCap / 2, "tile must fit the candidate queue twice");

inline bool skew_supported(int M, int d, int k) {
    (void)d;
    return M == 16 && k <= B200_IVFPQ_MAX_K;
}

__host__ __device__ inline size_t skew_smem_bytes(int d, int k) {
    return kSkewLutBytes + sizeof(float) * static_cast<size_t>((d + 3) & ~3) + TopK::smem_bytes(k, kScanCap) + 16;
}

// pq (M, 256, dsub) -> pq_t (256, dsub, M): the LUT build reads it with m fastest (coalesced).
__global__ void pq_transpose_kernel(const float* __restrict__ pq, float* __restrict__ pq_t, int M, int dsub) {
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    int64_t total = static_cast<int64_t>(M) * 256 * dsub;
    if (i >= total) return;
    int j = static_cast<int>(i % dsub);
    int c = static_cast<int>((i / dsub) % 256);
    int m = static_cast<int>(i / (static_cast<int64_t>(dsub) * 256));
    pq_t[(static_cast<int64_t>(c) * dsub + j) * M + m] = pq[i];
}

__device__ __forceinline__ uint4 skew_load_code16(const uint8_t* __restrict__ lcodes, int64_t idx, int64_t n) {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (idx < n) v = __ldg(reinterpret_cast<const uint4*>(lcodes) + idx);
    return v;
}

template <int B>
__device__ __forceinline__ float skew_lookup(const char* __restrict__ lutb, uint32_t w, uint32_t loff, int p) {
    // result byte0 = loff (4*lane < 128), byte1 = byte B of w, bytes 2,3 = 0  ->  (c << 8) | 4*lane
    uint32_t a = __byte_perm(w, loff, 0x6504 | (B << 4));
    return *reinterpret_cast<const float*>(lutb + a + 4 * p);
}

__global__ void __launch_bounds__(kThreads, 2) scan_skew16_kernel(const ScanParams p, const float* __restrict__ pq_t) {
    constexpr int M = 16;
    extern __shared__ __align__(1024) unsigned char smem_skew[];
    float* lut = reinterpret_cast<float*>(smem_skew);
    float* res = lut + 256 * kSkewRowWords;
    TopK tk;
    tk.bind(res + ((p.d + 3) & ~3), p.k, kScanCap);
    int* s_work = tk.meta + 4;
    const char* lutb = reinterpret_cast<const char*>(lut);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kWarps = kThreads / 32;
    const int r = lane & (M - 1);
    const int pstart = (M - r) & (M - 1);   // step at which this lane starts a new code (m == 0)
    const int pend = M - 1 - r;             // step at which this lane finishes a code (m == 15)
    float keep[M], cap[M];
#pragma unroll
    for (int s = 0; s < M; s++) {
        keep[s] = (s == pstart) ? 0.0f : 1.0f;
        cap[s] = (s == pend) ? 1.0f : 0.0f;
    }
    const uint32_t loff = static_cast<uint32_t>(lane) * 4u;
    const bool ws2 = (r & 8) != 0, ws1 = (r & 4) != 0;
    const uint32_t bs = static_cast<uint32_t>(r & 3) * 8u;
    const int nvalid = p.stats->nvalid;

    for (;;) {
        if (tid == 0) *s_work = atomicAdd(&p.stats->work_counter, 1);
        __syncthreads();
        const int wk = *s_work;
        if (wk >= nvalid) break;
        const int pair = p.order[wk];
        const int q = pair / p.nprobe;
        const int list = p.probe[pair];
        const int64_t beg = p.offsets[list];
        const int64_t n = p.offsets[list + 1] - beg;
        const uint8_t* lcodes = p.codes + beg * M;

        // issue the first code loads before building the LUT so that they overlap it
        const int64_t nblocks = (n + 31) >> 5;
        const int Bw = nblocks > warp ? static_cast<int>((nblocks - warp + kWarps - 1) / kWarps) : 0;
        const int maxBw = static_cast<int>((nblocks + kWarps - 1) / kWarps);
        auto code_index = [&](int b) -> int64_t {   // global code index of this lane in its b-th block
            return (static_cast<int64_t>(b) * kWarps + warp) * 32 + lane;
        };
        uint4 cur = make_uint4(0u, 0u, 0u, 0u);
        uint4 nxt = skew_load_code16(lcodes, 0 < Bw ? code_index(0) : n, n);
        uint4 pf1 = skew_load_code16(lcodes, 1 < Bw ? code_index(1) : n, n);

        // a2: residual
        for (int j = tid; j < p.d; j += kThreads)
            res[j] = __fsub_rn(p.xq[static_cast<int64_t>(q) * p.d + j], p.cent[static_cast<int64_t>(list) * p.d + j]);
        const uint32_t ext_thr = *reinterpret_cast<volatile uint32_t*>(p.qthr + q);
        if (tid == 0) tk.reset(ext_thr);
        __syncthreads();
        // a3: LUT, periodic rows.  idx -> (c = idx / 16, m = idx % 16): a warp writes 16 distinct banks
        // twice; pq_t is read with m fastest.
        for (int idx = tid; idx < M * 256; idx += kThreads) {
            const int m = idx & (M - 1), c = idx >> 4;
            const float* pc = pq_t + static_cast<int64_t>(c) * p.dsub * M + m;
            const float* rr = res + m * p.dsub;
            float acc = 0.0f;
            for (int j = 0; j < p.dsub; j++) acc = sqdiff_acc(acc, rr[j], __ldg(pc + j * M));
            float* row = lut + c * kSkewRowWords + m;
            row[0] = acc;
            row[16] = acc;
            row[32] = acc;
        }
        __syncthreads();

        // a4 + a5
        uint32_t thr = ext_thr;
        float acc = 0.0f;
        const int niter = maxBw + 1;   // +1: prologue block that only feeds bytes 0..r-1 of code 0
        for (int t0 = 0; t0 < niter; t0 += kSkewTB) {
#pragma unroll 1
            for (int tb = 0; tb < kSkewTB; tb++) {
                const int b = t0 + tb - 1;
                if (b < Bw) {
                    uint4 pf2 = skew_load_code16(lcodes, b + 3 < Bw ? code_index(b + 3) : n, n);
                    // window = bytes [r, r+16) of cur|nxt: word shift by r/4 (two select stages), then a
                    // funnel shift by 8*(r%4)
                    uint32_t x0 = cur.x, x1 = cur.y, x2 = cur.z, x3 = cur.w, x4 = nxt.x, x5 = nxt.y, x6 = nxt.z,
                             x7 = nxt.w;
                    uint32_t y0 = ws2 ? x2 : x0, y1 = ws2 ? x3 : x1, y2 = ws2 ? x4 : x2, y3 = ws2 ? x5 : x3,
                             y4 = ws2 ? x6 : x4, y5 = ws2 ? x7 : x5;
                    uint32_t z0 = ws1 ? y1 : y0, z1 = ws1 ? y2 : y1, z2 = ws1 ? y3 : y2, z3 = ws1 ? y4 : y3,
                             z4 = ws1 ? y5 : y4;
                    const uint32_t w0 = __funnelshift_r(z0, z1, bs), w1 = __funnelshift_r(z1, z2, bs),
                                   w2 = __funnelshift_r(z2, z3, bs), w3 = __funnelshift_r(z3, z4, bs);
                    float fin = 0.0f;
#define SKEW_STEP(W, B, P)                                             \
    {                                                                  \
        float T = skew_lookup<B>(lutb, W, loff, P);                    \
        acc = __fmaf_rn(acc, keep[P], T);                              \
        fin = __fmaf_rn(acc, cap[P], fin);                             \
    }
                    SKEW_STEP(w0, 0, 0) SKEW_STEP(w0, 1, 1) SKEW_STEP(w0, 2, 2) SKEW_STEP(w0, 3, 3)
                    SKEW_STEP(w1, 0, 4) SKEW_STEP(w1, 1, 5) SKEW_STEP(w1, 2, 6) SKEW_STEP(w1, 3, 7)
                    SKEW_STEP(w2, 0, 8) SKEW_STEP(w2, 1, 9) SKEW_STEP(w2, 2, 10) SKEW_STEP(w2, 3, 11)
                    SKEW_STEP(w3, 0, 12) SKEW_STEP(w3, 1, 13) SKEW_STEP(w3, 2, 14) SKEW_STEP(w3, 3, 15)
#undef SKEW_STEP
                    if (b >= 0) {
                        const int64_t idx = code_index(b);
                        const uint32_t bits = __float_as_uint(fin);
                        tk.push(idx < n && bits <= thr, make_key(bits, static_cast<uint32_t>(idx)));
                    }
                    cur = nxt;
                    nxt = pf1;
                    pf1 = pf2;
                }
            }
            tk.sync_and_flush_if_over<kThreads>(kScanCap - kThreads * kSkewTB, ext_thr);
            thr = tk.threshold();
        }
        __syncthreads();
        tk.flush<kThreads>(ext_thr);
        const int nb = tk.count();
        const uint64_t* s = tk.sorted();
        for (int i = tid; i < nb; i += kThreads) p.out_keys[static_cast<int64_t>(pair) * p.k + i] = s[i];
        if (tid == 0) {
            p.out_cnt[pair] = nb;
            if (nb == p.k) atomicMin(p.qthr + q, static_cast<uint32_t>(s[p.k - 1] >> 32));
        }
        __syncthreads();
    }
}

// returns 0, or -1 on a launch error (caller reads cudaGetLastError)
inline int launch_scan_skew(const ScanParams& sp, const float* pq_t, int64_t npairs, int num_sms, cudaStream_t st) {
    size_t smem = skew_smem_bytes(sp.d, sp.k);
    if (cudaFuncSetAttribute(scan_skew16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return -1;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scan_skew16_kernel, kThreads, smem) != cudaSuccess)
        return -1;
    if (per_sm < 1) return -1;
    int64_t grid = static_cast<int64_t>(per_sm) * num_sms;
    if (grid > npairs) grid = npairs;
    scan_skew16_kernel<<<(unsigned)grid, kThreads, smem, st>>>(sp, pq_t);
    return cudaPeekAtLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace b200
