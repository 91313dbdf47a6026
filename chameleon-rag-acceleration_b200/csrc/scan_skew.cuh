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
constexpr int kSkewCap = 2 * kThreads * kSkewTB;          // candidate queue: two tiles

inline bool skew_supported(int M, int d, int k) {
    (void)d;
    return M == 16 && k <= B200_IVFPQ_MAX_K;
}

__host__ __device__ inline size_t skew_smem_bytes(int d, int k) {
    return kSkewLutBytes + sizeof(float) * static_cast<size_t>((d + 3) & ~3) + TopK::smem_bytes(k, kSkewCap) + 16;
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

__device__ __forceinline__ uint4 skew_load_code16(const uint4* __restrict__ lp, uint32_t idx, uint32_t n) {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (idx < n) v = __ldg(lp + idx);
    return v;
}

template <int B>
__device__ __forceinline__ float skew_lookup(const char* __restrict__ lutb, uint32_t w, uint32_t loff, int p) {
    // result byte0 = loff (4*lane < 128), byte1 = byte B of w, bytes 2,3 = 0  ->  (c << 8) | 4*lane
    uint32_t a = __byte_perm(w, loff, 0x6504 | (B << 4));
    return *reinterpret_cast<const float*>(lutb + a + 4 * p);
}

// ---- LUT entry in packed fp32x2 arithmetic ------------------------------------------------------------------------
// sub.rn.f32x2 / mul.rn.f32x2 round each half exactly like the scalar ops, so (r_j - p_j)^2 is bit-identical to
// the oracle's; the accumulation stays scalar and sequential in j (FADD2 + FMUL2 + 2 FADD per two dimensions
// instead of 6 scalar ops).  The parity tests compare every LUT-dependent distance bit for bit, which also guards
// against the assembler contracting any of this into FMAs.
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
    return (static_cast<uint64_t>(__float_as_uint(hi)) << 32) | __float_as_uint(lo);
}
__device__ __forceinline__ uint64_t sub_f32x2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t mul_f32x2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// T = sum_j (r_j - p_j)^2 for one (m, c) entry; p points at pq_t[c][0][m] (stride M between dimensions, so that the
// 16 lanes with consecutive m read consecutive words), rr2 holds (r_j, r_j+1)
template <int DSUB, int M>
__device__ __forceinline__ float lut_entry_packed(const float* __restrict__ p, const uint64_t (&rr2)[DSUB / 2]) {
    float pv[DSUB];
#pragma unroll
    for (int j = 0; j < DSUB; j++) pv[j] = __ldg(p + j * M);
    float a = 0.0f;
#pragma unroll
    for (int j = 0; j < DSUB / 2; j++) {
        const uint64_t d2 = sub_f32x2(rr2[j], pack_f32x2(pv[2 * j], pv[2 * j + 1]));
        const uint64_t s2 = mul_f32x2(d2, d2);
        a = __fadd_rn(a, __uint_as_float(static_cast<uint32_t>(s2)));
        a = __fadd_rn(a, __uint_as_float(static_cast<uint32_t>(s2 >> 32)));
    }
    return a;
}

// One block = 16 lookups per lane over the byte window [r, r+16) of cur|nxt.  acc carries the partial sum of the
// code the lane is on; the return value is the finished distance of the code that completed in this block.
__device__ __forceinline__ float skew_block16(const char* __restrict__ lutb, const uint4& cur, const uint4& nxt,
                                              bool ws2, bool ws1, uint32_t bs, uint32_t loff,
                                              const float (&keep)[16], const float (&cap)[16], float& acc) {
    // window = bytes [r, r+16) of cur|nxt: word shift by r/4 (two select stages), then a funnel shift by 8*(r%4)
    const uint32_t y0 = ws2 ? cur.z : cur.x, y1 = ws2 ? cur.w : cur.y, y2 = ws2 ? nxt.x : cur.z,
                   y3 = ws2 ? nxt.y : cur.w, y4 = ws2 ? nxt.z : nxt.x, y5 = ws2 ? nxt.w : nxt.y;
    const uint32_t z0 = ws1 ? y1 : y0, z1 = ws1 ? y2 : y1, z2 = ws1 ? y3 : y2, z3 = ws1 ? y4 : y3,
                   z4 = ws1 ? y5 : y4;
    const uint32_t w0 = __funnelshift_r(z0, z1, bs), w1 = __funnelshift_r(z1, z2, bs),
                   w2 = __funnelshift_r(z2, z3, bs), w3 = __funnelshift_r(z3, z4, bs);
    float fin = 0.0f;
#define SKEW_STEP(W, B, P)                              \
    {                                                   \
        float T = skew_lookup<B>(lutb, W, loff, P);     \
        acc = __fmaf_rn(acc, keep[P], T);               \
        fin = __fmaf_rn(acc, cap[P], fin);              \
    }
    SKEW_STEP(w0, 0, 0) SKEW_STEP(w0, 1, 1) SKEW_STEP(w0, 2, 2) SKEW_STEP(w0, 3, 3)
    SKEW_STEP(w1, 0, 4) SKEW_STEP(w1, 1, 5) SKEW_STEP(w1, 2, 6) SKEW_STEP(w1, 3, 7)
    SKEW_STEP(w2, 0, 8) SKEW_STEP(w2, 1, 9) SKEW_STEP(w2, 2, 10) SKEW_STEP(w2, 3, 11)
    SKEW_STEP(w3, 0, 12) SKEW_STEP(w3, 1, 13) SKEW_STEP(w3, 2, 14) SKEW_STEP(w3, 3, 15)
#undef SKEW_STEP
    return fin;
}

// DSUB = d / 16 when it is one of the specialised values (residual slice held in registers), 0 = generic.
template <int DSUB>
__global__ void __launch_bounds__(kThreads, 2) scan_skew16_kernel(const ScanParams p, const float* __restrict__ pq_t) {
    constexpr int M = 16;
    extern __shared__ __align__(1024) unsigned char smem_skew[];
    float* lut = reinterpret_cast<float*>(smem_skew);
    float* res = lut + 256 * kSkewRowWords;
    TopK tk;
    tk.bind(res + ((p.d + 3) & ~3), p.k, kSkewCap);
    int* s_work = tk.meta + 4;
    const char* lutb = reinterpret_cast<const char*>(lut);

    const int tid = threadIdx.x, lane = tid & 31;
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
    const int dsub = DSUB ? DSUB : p.dsub;
    // LUT build mapping: this thread owns sub-quantizer lm and code values lc0 + 16 i; the two half-warps store
    // their 4 periodic copies in opposite order so that the 32 lanes always hit 32 distinct banks
    const int lm = tid & (M - 1), lc0 = tid >> 4;
    const int copy0 = (lane >> 4) * 16;

    // work items are pulled from a global counter; the fetch for the NEXT item is issued at the top of the current
    // one so that the atomic's round trip to L2 overlaps the pair's work
    int next_work = 0;
    if (tid == 0) next_work = atomicAdd(&p.stats->work_counter, 1);
    for (;;) {
        if (tid == 0) *s_work = next_work;
        __syncthreads();
        const int wk = *s_work;
        if (wk >= nvalid * p.nseg) break;
        if (tid == 0) next_work = atomicAdd(&p.stats->work_counter, 1);
        const int pair = p.order[wk / p.nseg];
        const int seg = wk % p.nseg;
        const int slot = pair * p.nseg + seg;
        const int q = pair / p.nprobe;
        const int list = p.probe[pair];
        const int64_t beg = p.offsets[list];
        const uint32_t ntot = static_cast<uint32_t>(p.offsets[list + 1] - beg);
        const uint32_t seglen = (((ntot + p.nseg - 1) / p.nseg) + 255u) & ~255u;
        const uint32_t soff = seg * seglen;                           // first code of this segment
        if (soff >= ntot) {                                           // empty segment: out_cnt stays 0
            __syncthreads();
            continue;
        }
        const uint32_t n = min(seglen, ntot - soff);
        const uint4* lp = reinterpret_cast<const uint4*>(p.codes + (beg + soff) * M);

        // lane's code in block b is b*256 + tid.  Ring of four code registers: slot (b+1)%4 holds code b.
        // Issue the first loads before building the LUT so that they overlap it.
        uint4 c0 = make_uint4(0u, 0u, 0u, 0u);                    // "code -1" of the prologue block
        uint4 c1 = skew_load_code16(lp, tid, n);                  // code 0
        uint4 c2 = skew_load_code16(lp, 256u + tid, n);           // code 1
        uint4 c3;

        const uint32_t ext_thr = *reinterpret_cast<volatile uint32_t*>(p.qthr + q);
        if (tid == 0) tk.reset(ext_thr);
        if (p.lutg) {
            // small batches: the pair's LUT was built once by lut_small_kernel, [c][m] layout (a half-warp reads the 16
            // entries of one code value: 64 contiguous bytes); only the periodic copies are made here
            const float* src = p.lutg + static_cast<int64_t>(pair) * (M * 256);
#pragma unroll 4
            for (int i = 0; i < 16; i++) {
                const int c = lc0 + 16 * i;
                const float a = __ldg(src + c * M + lm);
                float* row = lut + c * kSkewRowWords + lm;
                row[copy0] = a;
                row[copy0 ^ 16] = a;
                row[32 + copy0] = a;
                row[32 + (copy0 ^ 16)] = a;
            }
        } else {
        // a2: residual
        for (int j = tid; j < p.d; j += kThreads)
            res[j] = __fsub_rn(p.xq[static_cast<int64_t>(q) * p.d + j], p.cent[static_cast<int64_t>(list) * p.d + j]);
        __syncthreads();
        // a3: LUT with periodic rows lut[c][w] = T[w % 16][c], w < 64
        if constexpr (DSUB != 0) {
            uint64_t rr2[DSUB / 2];
#pragma unroll
            for (int j = 0; j < DSUB / 2; j++) rr2[j] = pack_f32x2(res[lm * DSUB + 2 * j], res[lm * DSUB + 2 * j + 1]);
#pragma unroll 4
            for (int i = 0; i < 16; i++) {
                const int c = lc0 + 16 * i;
                const float a = lut_entry_packed<DSUB, M>(pq_t + static_cast<int64_t>(c) * (DSUB * M) + lm, rr2);
                float* row = lut + c * kSkewRowWords + lm;
                row[copy0] = a;
                row[copy0 ^ 16] = a;
                row[32 + copy0] = a;
                row[32 + (copy0 ^ 16)] = a;
            }
        } else {
            for (int i = 0; i < 16; i++) {
                const int c = lc0 + 16 * i;
                const float* pc = pq_t + static_cast<int64_t>(c) * dsub * M + lm;
                const float* rr = res + lm * dsub;
                float a = 0.0f;
                for (int j = 0; j < dsub; j++) a = sqdiff_acc(a, rr[j], __ldg(pc + j * M));
                float* row = lut + c * kSkewRowWords + lm;
                row[copy0] = a;
                row[copy0 ^ 16] = a;
                row[32 + copy0] = a;
                row[32 + (copy0 ^ 16)] = a;
            }
        }
        }
        __syncthreads();

        // a4 + a5.  Iteration `it` processes block b = it - 1 (it = 0 is the prologue that only feeds bytes
        // 0..r-1 of code 0).  Four iterations per tile, fully unrolled so that the code ring needs no moves.
        uint32_t thr = ext_thr;
        float acc = 0.0f;
        const uint32_t nblk = (n + 255u) >> 8;
        for (uint32_t t0 = 0; t0 <= nblk; t0 += kSkewTB) {
            uint32_t base = t0 * 256u + tid;            // code index of block b = t0 (iteration t0 + 1)
#define SKEW_ITER(CUR, NXT, LOADTO, TB)                                                         \
    {                                                                                           \
        LOADTO = skew_load_code16(lp, base + (TB + 2) * 256u, n);                               \
        const float fin = skew_block16(lutb, CUR, NXT, ws2, ws1, bs, loff, keep, cap, acc);     \
        const uint32_t idx = base + TB * 256u - 256u; /* wraps past 2^32 for the prologue block */ \
        const uint32_t bits = __float_as_uint(fin);                                             \
        tk.push(idx < n && bits <= thr, make_key(bits, soff + idx));                                   \
    }
#pragma unroll
            for (int half = 0; half < kSkewTB / 4; half++) {
                SKEW_ITER(c0, c1, c3, 0)
                SKEW_ITER(c1, c2, c0, 1)
                SKEW_ITER(c2, c3, c1, 2)
                SKEW_ITER(c3, c0, c2, 3)
                base += 1024u;
            }
#undef SKEW_ITER
            // no threshold yet (cold start): fold the first tile in right away so that later tiles are filtered
            tk.sync_and_flush_if_over<kThreads>(thr == kInfBits ? 0 : kSkewCap - kThreads * kSkewTB, ext_thr);
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
        __syncthreads();
    }
}

template <int DSUB>
int launch_scan_skew_t(const ScanParams& sp, const float* pq_t, int64_t npairs, int num_sms, cudaStream_t st) {
    size_t smem = skew_smem_bytes(sp.d, sp.k);
    auto kernel = scan_skew16_kernel<DSUB>;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem) != cudaSuccess) return -1;
    if (per_sm < 1) return -1;
    int64_t grid = static_cast<int64_t>(per_sm) * num_sms;
    if (grid > npairs * sp.nseg) grid = npairs * sp.nseg;
    kernel<<<(unsigned)grid, kThreads, smem, st>>>(sp, pq_t);
    return cudaPeekAtLastError() == cudaSuccess ? 0 : -1;
}

// returns 0, or -1 on a launch error (caller reads cudaGetLastError)
inline int launch_scan_skew(const ScanParams& sp, const float* pq_t, int64_t npairs, int num_sms, cudaStream_t st) {
    switch (sp.dsub) {
        case 4: return launch_scan_skew_t<4>(sp, pq_t, npairs, num_sms, st);
        case 6: return launch_scan_skew_t<6>(sp, pq_t, npairs, num_sms, st);
        case 8: return launch_scan_skew_t<8>(sp, pq_t, npairs, num_sms, st);
        case 16: return launch_scan_skew_t<16>(sp, pq_t, npairs, num_sms, st);
        default: return launch_scan_skew_t<0>(sp, pq_t, npairs, num_sms, st);
    }
}

}  // namespace b200
