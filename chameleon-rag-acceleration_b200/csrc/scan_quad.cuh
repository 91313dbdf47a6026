// scan_quad.cuh -- M = 16 ADC scan for FOUR queries of the same list: an integer lower-bound FILTER in front of the
// exact fp32 evaluation.  (K1's idea -- cheap provable pre-filter, exact rescoring of the survivors -- applied to K3.)
//
// Why.  The exact scans (scan_skew.cuh, scan_duo.cuh) are bound by the shared-memory data pipe: one fp32 table entry
// (4 B) per code byte and query, 128 B per clock and SM.  An fp32 sum also has to be added in the oracle's order,
// which costs the skewed kernels a restart/capture FMA pair per step.  Almost all of that work is spent on codes that
// are nowhere near the current k-th best distance.
//
// How.
//   * Next to the exact LUT T_q[m][c] (fp32, kept in global memory / L2, one float4 per (m, c) for the four queries)
//     the LUT build writes an 11-bit quantised copy to shared memory:  u_q[m][c] = floor(T_q[m][c] * s_q)  <= 2047,
//     s_q = 2047 / B_q with B_q >= every entry of query q ((||r_m|| + max_c ||p_mc||)^2, triangle inequality).
//     Four queries share one 8-byte entry (4 x u16); rows are indexed by the code value, 32 periodic entries per
//     256-byte row exactly like scan_duo.cuh, so ONE conflict-free LDS.64 fetches the entry of four queries.
//   * Integer sums are exact in any order, so a lane can walk the 16 bytes of ITS OWN code in rotated order
//     (m = r, r+1, .., 15, 0, .., r-1 with r = lane % 16): the bank-conflict-free skew needs no restart/capture any
//     more.  Two packed 32-bit accumulators (u16 x 2 each, sums <= 16 x 2047 < 2^15: no carry between the halves)
//     hold the four lower bounds: per step PRMT + LDS.64 + 2 IADD for FOUR look-ups.
//   * LB_q(code) = sum_m u_q[m][code_m] satisfies LB_q <= s_q * exact_q (1 + 1.1e-6), so a code can only be among
//     the results if LB_q <= t_q = floor(thr_q * s_q * (1 + 4e-6)) + 1, thr_q being the query's current k-th best
//     distance (checked for the four queries with two subtractions on guard-bit packed words).  Everything else is
//     dropped without ever touching fp32.
//   * Survivors (a few per thousand codes once a threshold exists) are queued as (offset, query mask); at tile
//     boundaries they are evaluated EXACTLY: 16 sequential fp32 adds of the exact LUT entries in ascending m --
//     bit-identical to the oracle -- and pushed into the same top-k machinery as the other kernels.
// The result set is therefore exactly the oracle's; only the amount of exact work changes.
//
// Reference semantics: ADC.hpp:75-99 / IVFPQ_1B_search.ipynb:7948-7960 (sum over m ascending),
// LUT_construction.hpp:180-209 / ipynb:7929-7946 (LUT), priority_queue_L1.hpp:65-75 (strict <).
#pragma once
#include "scan_duo.cuh"

namespace b200 {

constexpr int kQuadLutBytes = 256 * 256;          // one table: 256 code values x 32 periodic entries x 4 x u16
constexpr int kQuadCap = 1024;                    // exact-candidate queue per query (M = 16)
constexpr int kQuadSurvMin = 1280;                // survivor queue (u32 each): one tile of 1024 codes + the drain trigger
constexpr int kQuadSurvMax = 1280;
constexpr uint32_t kQuadMaxList = 1u << 28;       // survivor entry = (offset << 4) | query mask

// M = 16: 256 threads, two CTAs per SM, one table.  M = 32: the 32-byte code is two independent 16-byte halves (integer
// sums do not care about the order), each with its own table: 512 threads, one CTA per SM, entries quantised to 10
// bits so that 32 of them still sum below 2^15.
template <int M>
struct QuadCfg {
    static_assert(M == 16 || M == 32, "M = 16 or 32");
    static constexpr int kT = 16 * M;                       // threads per CTA
    static constexpr int kTables = M / 16;
    static constexpr int kCtasPerSm = M == 16 ? 2 : 1;
    static constexpr uint32_t kQMax = M == 16 ? 2047u : 1023u;
    static constexpr int kCap = M == 16 ? kQuadCap : 2048;  // candidate queue per query
    static constexpr int kTileBlocks = 1024 / kT;           // blocks of kT codes between survivor checks
    static constexpr size_t kScratchFloat4 = static_cast<size_t>(M) * 256;   // exact LUT of one CTA: [m][c] float4
};
inline size_t quad_scratch_float4(int M) { return static_cast<size_t>(M) * 256; }

inline bool quad_supported(int M, int d, int k) {
    (void)d;
    return (M == 16 || M == 32) && k <= 512;
}

// shared memory: [ lut16 | residuals (4 x dpad f32) | 4 x TopK | survivors | control ]
__host__ __device__ inline size_t quad_smem_fixed(int M, int d, int k) {
    return static_cast<size_t>(M / 16) * kQuadLutBytes + 4 * sizeof(float) * static_cast<size_t>((d + 3) & ~3) +
           4 * TopK::smem_bytes(k, M == 16 ? kQuadCap : 2048) + 64 + 2 * sizeof(QuadGroup);
}
// survivor queue capacity: whatever two CTAs per SM leave (228 KB per SM, 1 KB reserved per CTA), in steps of 256
__host__ __device__ inline int quad_surv_cap(int M, int d, int k) {
    const long long room = (228 * 1024 / 2 - 1024 - 256) - static_cast<long long>(quad_smem_fixed(M, d, k));
    long long cap = room > 0 ? (room / 4) & ~255ll : 0;
    if (cap > kQuadSurvMax) cap = kQuadSurvMax;
    if (cap < kQuadSurvMin) cap = kQuadSurvMin;
    return static_cast<int>(cap);
}
__host__ __device__ inline size_t quad_smem_bytes(int M, int d, int k) {
    return quad_smem_fixed(M, d, k) + sizeof(uint32_t) * quad_surv_cap(M, d, k);
}

struct QuadCtrl {          // 64 bytes
    int work;              // current work item
    int nsurv[2];          // survivors queued; the two counters alternate from one drain to the next, so that the one
                           // in use was zeroed a whole drain (several barriers) ago
    int pad0_;
    float scale[4];        // s_q (generic path only)
    uint32_t pad_[8];
};
static_assert(sizeof(QuadCtrl) == 64, "QuadCtrl is 64 bytes");

// max_c ||pq[m][c]|| per sub-quantizer, rounded up a little: one CTA of 256 threads (one per code value) per m
__global__ void pq_maxnorm_kernel(const float* __restrict__ pq, int dsub, float* __restrict__ out) {
    __shared__ float s[256];
    const int m = blockIdx.x, c = threadIdx.x;
    const float* pc = pq + (static_cast<int64_t>(m) * 256 + c) * dsub;
    float a = 0.0f;
    for (int j = 0; j < dsub; j++) a += pc[j] * pc[j];
    s[c] = a;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (c < o) s[c] = fmaxf(s[c], s[c + o]);
        __syncthreads();
    }
    if (c == 0) out[m] = sqrtf(s[0]) * 1.00001f;
}

__device__ __forceinline__ void quad_copy_group_async(QuadGroup* dst, const QuadGroup* src) {
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
    const char* s = reinterpret_cast<const char*>(src);
#pragma unroll
    for (int i = 0; i < 3; i++)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d + 16 * i), "l"(s + 16 * i) : "memory");
}

// integer threshold of one query: everything with LB <= t may still have an exact distance <= thr
__device__ __forceinline__ uint32_t quad_int_threshold(uint32_t thr_bits, float scale) {
    if (thr_bits >= kInfBits) return 0x7fffu;
    const float x = __uint_as_float(thr_bits) * scale * 1.000004f;
    if (!(x < 32000.0f)) return 0x7fffu;
    return static_cast<uint32_t>(x) + 1u;
}

template <int B>
__device__ __forceinline__ uint2 quad_lookup(const char* __restrict__ lutb, uint32_t w, uint32_t loff, int p) {
    uint32_t a = __byte_perm(w, loff, 0x6504 | (B << 4));
    return *reinterpret_cast<const uint2*>(lutb + a + 8 * p);
}

template <int M>
struct QuadCode {
    uint4 v[M / 16];
};
template <int M>
__device__ __forceinline__ QuadCode<M> quad_load_code(const uint4* __restrict__ lp, uint32_t idx, uint32_t n) {
    QuadCode<M> c;
#pragma unroll
    for (int h = 0; h < M / 16; h++) c.v[h] = make_uint4(0u, 0u, 0u, 0u);
    if (idx < n) {
#pragma unroll
        for (int h = 0; h < M / 16; h++) c.v[h] = __ldg(lp + static_cast<size_t>(idx) * (M / 16) + h);
    }
    return c;
}

// the lane's code walked in rotated byte order; returns the four packed lower bounds
__device__ __forceinline__ uint2 quad_block16(const char* __restrict__ lutb, const uint4& code, bool ws2, bool ws1,
                                              uint32_t bs, uint32_t loff) {
    // rotate the 16 bytes left by r: word rotation by r / 4, then a funnel shift by 8 * (r % 4) with wrap-around
    const uint32_t y0 = ws2 ? code.z : code.x, y1 = ws2 ? code.w : code.y, y2 = ws2 ? code.x : code.z,
                   y3 = ws2 ? code.y : code.w;
    const uint32_t z0 = ws1 ? y1 : y0, z1 = ws1 ? y2 : y1, z2 = ws1 ? y3 : y2, z3 = ws1 ? y0 : y3;
    const uint32_t w0 = __funnelshift_r(z0, z1, bs), w1 = __funnelshift_r(z1, z2, bs),
                   w2 = __funnelshift_r(z2, z3, bs), w3 = __funnelshift_r(z3, z0, bs);
    uint32_t s01 = 0u, s23 = 0u;
#define QUAD_STEP(W, B, P)                                 \
    {                                                      \
        const uint2 t = quad_lookup<B>(lutb, W, loff, P);  \
        s01 += t.x;                                        \
        s23 += t.y;                                        \
    }
    QUAD_STEP(w0, 0, 0) QUAD_STEP(w0, 1, 1) QUAD_STEP(w0, 2, 2) QUAD_STEP(w0, 3, 3)
    QUAD_STEP(w1, 0, 4) QUAD_STEP(w1, 1, 5) QUAD_STEP(w1, 2, 6) QUAD_STEP(w1, 3, 7)
    QUAD_STEP(w2, 0, 8) QUAD_STEP(w2, 1, 9) QUAD_STEP(w2, 2, 10) QUAD_STEP(w2, 3, 11)
    QUAD_STEP(w3, 0, 12) QUAD_STEP(w3, 1, 13) QUAD_STEP(w3, 2, 14) QUAD_STEP(w3, 3, 15)
#undef QUAD_STEP
    return make_uint2(s01, s23);
}

// DSUB = d / 16 when it is one of the specialised values (residuals held in registers), 0 = generic.
template <int M, int DSUB>
__global__ void __launch_bounds__(QuadCfg<M>::kT, QuadCfg<M>::kCtasPerSm)
scan_quad_kernel(const ScanParams p, const float* __restrict__ pq_t) {
    using Cfg = QuadCfg<M>;
    constexpr int kT = Cfg::kT;
    constexpr int kCap = Cfg::kCap;
    extern __shared__ __align__(1024) unsigned char smem_quad[];
    uint2* lut16 = reinterpret_cast<uint2*>(smem_quad);                           // kTables tables, 64 KB each
    const char* lutb = reinterpret_cast<const char*>(lut16);
    const int dpad = (p.d + 3) & ~3;
    float4* res4 = reinterpret_cast<float4*>(smem_quad + Cfg::kTables * kQuadLutBytes);   // [d] (r_0, r_1, r_2, r_3)
    unsigned char* tk_base = reinterpret_cast<unsigned char*>(res4 + dpad);
    TopK tk[4];
#pragma unroll
    for (int q = 0; q < 4; q++) tk[q].bind(tk_base + q * TopK::smem_bytes(p.k, kCap), p.k, kCap);
    uint32_t* surv = reinterpret_cast<uint32_t*>(tk_base + 4 * TopK::smem_bytes(p.k, kCap));
    const int surv_cap = quad_surv_cap(M, p.d, p.k);
    QuadCtrl* ctrl = reinterpret_cast<QuadCtrl*>(surv + surv_cap);
    QuadGroup* s_grp = reinterpret_cast<QuadGroup*>(ctrl + 1);
    float4* lutf = p.lutf_scratch + static_cast<size_t>(blockIdx.x) * Cfg::kScratchFloat4;   // exact LUT [m][c]

    const int tid = threadIdx.x, lane = tid & 31;
    const int r = lane & 15;
    const uint32_t loff = static_cast<uint32_t>(r) * 8u;
    const bool ws2 = (r & 8) != 0, ws1 = (r & 4) != 0;
    const uint32_t bs = static_cast<uint32_t>(r & 3) * 8u;
    const int ngroups = p.stats->ngroups;
    const int dsub = DSUB ? DSUB : p.dsub;
    const int lm = tid & (M - 1), lc0 = tid / M;   // LUT build: sub-quantizer lm, code values lc0 + 16 i

    int next_work = 0, buf = 0;
    if (tid == 0) {
        ctrl->nsurv[0] = 0;
        ctrl->nsurv[1] = 0;
        next_work = atomicAdd(&p.stats->work_counter, 1);
        if (next_work < ngroups) quad_copy_group_async(&s_grp[0], static_cast<const QuadGroup*>(p.groups) + next_work);
    }
    for (;;) {
        if (tid == 0) {
            ctrl->work = next_work;
            asm volatile("cp.async.wait_all;" ::: "memory");
        }
        __syncthreads();
        const int wk = ctrl->work;
        if (wk >= ngroups) break;
        if (tid == 0) next_work = atomicAdd(&p.stats->work_counter, 1);
        const QuadGroup grp = s_grp[buf];
        int pair[4], qi[4];
        uint32_t vmask = 0u;   // which of the four slots hold a query
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const bool has = grp.pair[q] >= 0;
            pair[q] = has ? grp.pair[q] : grp.pair[0];
            qi[q] = pair[q] / p.nprobe;
            vmask |= has ? (1u << q) : 0u;
        }
        const int list = grp.list;
        const uint32_t n = grp.n;
        const uint4* lp = reinterpret_cast<const uint4*>(p.codes + grp.beg * M);
        constexpr float kQ = static_cast<float>(Cfg::kQMax);

        uint32_t ext[4];
#pragma unroll
        for (int q = 0; q < 4; q++) ext[q] = *reinterpret_cast<volatile uint32_t*>(p.qthr + qi[q]);
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (tid == q) tk[q].reset(ext[q]);
        // a2: residuals + the quantisation scale s_q = 2047 / B_q, B_q = max_m (||r_m|| + max_c ||p_mc||)^2 >= every
        // table entry of query q.
        float s0, s1, s2, s3;
        uint64_t r01[DSUB ? DSUB : 1], r23[DSUB ? DSUB : 1];
        if constexpr (DSUB != 0) {
            // specialised path: every thread fetches the slice of ITS sub-quantizer lm straight from global memory
            // (no shared-memory staging, no barrier); the 16 lanes of a half-warp cover the 16 sub-quantizers, so the
            // max over m is four butterfly shuffles and every thread ends up with the same scales
            const float* cp = p.cent + static_cast<int64_t>(list) * p.d + lm * DSUB;
            const float* x0 = p.xq + static_cast<int64_t>(qi[0]) * p.d + lm * DSUB;
            const float* x1 = p.xq + static_cast<int64_t>(qi[1]) * p.d + lm * DSUB;
            const float* x2 = p.xq + static_cast<int64_t>(qi[2]) * p.d + lm * DSUB;
            const float* x3 = p.xq + static_cast<int64_t>(qi[3]) * p.d + lm * DSUB;
            float n0 = 0.0f, n1 = 0.0f, n2 = 0.0f, n3 = 0.0f;
#pragma unroll
            for (int j = 0; j < DSUB; j++) {
                const float cj = __ldg(cp + j);
                const float a0 = __fsub_rn(__ldg(x0 + j), cj), a1 = __fsub_rn(__ldg(x1 + j), cj),
                            a2 = __fsub_rn(__ldg(x2 + j), cj), a3 = __fsub_rn(__ldg(x3 + j), cj);
                r01[j] = pack_f32x2(a0, a1);
                r23[j] = pack_f32x2(a2, a3);
                n0 += a0 * a0;
                n1 += a1 * a1;
                n2 += a2 * a2;
                n3 += a3 * a3;
            }
            const float pm = p.pq_maxnorm[lm];
            float b0 = sqrtf(n0) * 1.00001f + pm, b1 = sqrtf(n1) * 1.00001f + pm, b2 = sqrtf(n2) * 1.00001f + pm,
                  b3 = sqrtf(n3) * 1.00001f + pm;
            b0 = b0 * b0 * 1.0001f;
            b1 = b1 * b1 * 1.0001f;
            b2 = b2 * b2 * 1.0001f;
            b3 = b3 * b3 * 1.0001f;
#pragma unroll
            for (int o = M / 2; o > 0; o >>= 1) {
                b0 = fmaxf(b0, __shfl_xor_sync(0xffffffffu, b0, o));
                b1 = fmaxf(b1, __shfl_xor_sync(0xffffffffu, b1, o));
                b2 = fmaxf(b2, __shfl_xor_sync(0xffffffffu, b2, o));
                b3 = fmaxf(b3, __shfl_xor_sync(0xffffffffu, b3, o));
            }
            s0 = b0 > 0.0f ? (kQ / b0) * 0.999999f : 0.0f;
            s1 = b1 > 0.0f ? (kQ / b1) * 0.999999f : 0.0f;
            s2 = b2 > 0.0f ? (kQ / b2) * 0.999999f : 0.0f;
            s3 = b3 > 0.0f ? (kQ / b3) * 0.999999f : 0.0f;
        } else {
            for (int j = tid; j < p.d; j += kT) {
                const float cj = p.cent[static_cast<int64_t>(list) * p.d + j];
                float4 rr;
                rr.x = __fsub_rn(p.xq[static_cast<int64_t>(qi[0]) * p.d + j], cj);
                rr.y = __fsub_rn(p.xq[static_cast<int64_t>(qi[1]) * p.d + j], cj);
                rr.z = __fsub_rn(p.xq[static_cast<int64_t>(qi[2]) * p.d + j], cj);
                rr.w = __fsub_rn(p.xq[static_cast<int64_t>(qi[3]) * p.d + j], cj);
                res4[j] = rr;
            }
            __syncthreads();
            if (tid < 4 * M) {
                const int q = tid / M, m = tid % M;
                float a = 0.0f;
                for (int j = 0; j < dsub; j++) {
                    const float4 rr = res4[m * dsub + j];
                    const float v = q == 0 ? rr.x : q == 1 ? rr.y : q == 2 ? rr.z : rr.w;
                    a += v * v;
                }
                float b = sqrtf(a) * 1.00001f + p.pq_maxnorm[m];
                b = b * b * 1.0001f;
                for (int o = M / 2; o > 0; o >>= 1) b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
                if (m == 0) ctrl->scale[q] = b > 0.0f ? (kQ / b) * 0.999999f : 0.0f;
            }
            __syncthreads();
            s0 = ctrl->scale[0];
            s1 = ctrl->scale[1];
            s2 = ctrl->scale[2];
            s3 = ctrl->scale[3];
        }
        // a3: exact LUT (global, float4 per (m, c)) + quantised copy (shared, periodic rows)
#define QUAD_LUT_STORE(C, T0, T1, T2, T3)                                                          \
    {                                                                                              \
        lutf[lm * 256 + (C)] = make_float4(T0, T1, T2, T3);                                        \
        const uint32_t u0 = min(static_cast<uint32_t>(__float2uint_rz((T0) * s0)), Cfg::kQMax),    \
                       u1 = min(static_cast<uint32_t>(__float2uint_rz((T1) * s1)), Cfg::kQMax),    \
                       u2 = min(static_cast<uint32_t>(__float2uint_rz((T2) * s2)), Cfg::kQMax),    \
                       u3 = min(static_cast<uint32_t>(__float2uint_rz((T3) * s3)), Cfg::kQMax);    \
        const uint2 e_ = make_uint2(u0 | (u1 << 16), u2 | (u3 << 16));                             \
        /* table lm / 16, row = code value, periodic entries (lm % 16) and (lm % 16) + 16 */       \
        uint2* row_ = lut16 + (lm >> 4) * (kQuadLutBytes / 8) + (C) * 32 + (lm & 15);              \
        row_[0] = e_;                                                                              \
        row_[16] = e_;                                                                             \
    }
        if constexpr (DSUB != 0) {
            const uint64_t negzero2 = p.negzero2;
            // PQ centroid slices are loaded two entries ahead of their use
#define QUAD_LUT_LOAD(PV, I)                                                                       \
    {                                                                                              \
        const float* pc_ = pq_t + static_cast<int64_t>(lc0 + 16 * (I)) * (DSUB * M) + lm;          \
        _Pragma("unroll") for (int j = 0; j < DSUB; j++) PV[j] = __ldg(pc_ + j * M);               \
    }
#define QUAD_LUT_EMIT(PV, I)                                                                       \
    {                                                                                              \
        const uint64_t e01 = lut_entry_duo<DSUB>(PV, r01, negzero2);                               \
        const uint64_t e23 = lut_entry_duo<DSUB>(PV, r23, negzero2);                               \
        const float T0 = __uint_as_float(static_cast<uint32_t>(e01)),                              \
                    T1 = __uint_as_float(static_cast<uint32_t>(e01 >> 32)),                        \
                    T2 = __uint_as_float(static_cast<uint32_t>(e23)),                              \
                    T3 = __uint_as_float(static_cast<uint32_t>(e23 >> 32));                        \
        QUAD_LUT_STORE(lc0 + 16 * (I), T0, T1, T2, T3)                                             \
    }
            float pv0[DSUB], pv1[DSUB], pv2[DSUB], pv3[DSUB];
            QUAD_LUT_LOAD(pv0, 0)
            QUAD_LUT_LOAD(pv1, 1)
#pragma unroll 1
            for (int i = 0; i < 16; i += 4) {
                QUAD_LUT_LOAD(pv2, i + 2)
                QUAD_LUT_LOAD(pv3, i + 3)
                QUAD_LUT_EMIT(pv0, i)
                QUAD_LUT_EMIT(pv1, i + 1)
                if (i + 4 < 16) {
                    QUAD_LUT_LOAD(pv0, i + 4)
                    QUAD_LUT_LOAD(pv1, i + 5)
                }
                QUAD_LUT_EMIT(pv2, i + 2)
                QUAD_LUT_EMIT(pv3, i + 3)
            }
#undef QUAD_LUT_LOAD
#undef QUAD_LUT_EMIT
        } else {
            for (int i = 0; i < 16; i++) {
                const int c = lc0 + 16 * i;
                const float* pc = pq_t + static_cast<int64_t>(c) * dsub * M + lm;
                float T0 = 0.0f, T1 = 0.0f, T2 = 0.0f, T3 = 0.0f;
                for (int j = 0; j < dsub; j++) {
                    const float pj = __ldg(pc + j * M);
                    const float4 rr = res4[lm * dsub + j];
                    T0 = sqdiff_acc(T0, rr.x, pj);
                    T1 = sqdiff_acc(T1, rr.y, pj);
                    T2 = sqdiff_acc(T2, rr.z, pj);
                    T3 = sqdiff_acc(T3, rr.w, pj);
                }
                QUAD_LUT_STORE(c, T0, T1, T2, T3)
            }
        }
#undef QUAD_LUT_STORE
        __syncthreads();   // also makes the exact LUT (global) visible to the whole CTA
        buf ^= 1;
        if (tid == 0 && next_work < ngroups)
            quad_copy_group_async(&s_grp[buf], static_cast<const QuadGroup*>(p.groups) + next_work);

        // Exact thresholds (fp32 bits) and their integer images are kept per thread: every thread reads the same
        // shared-memory words after the same barriers, so the copies agree.
        uint32_t th0 = ext[0], th1 = ext[1], th2 = ext[2], th3 = ext[3];
        uint32_t t01, t23;
        auto refresh_int_thresholds = [&]() {
            const uint32_t i0 = quad_int_threshold(th0, s0), i1 = quad_int_threshold(th1, s1),
                           i2 = quad_int_threshold(th2, s2), i3 = quad_int_threshold(th3, s3);
            t01 = 0x80008000u | (i1 << 16) | i0;
            t23 = 0x80008000u | (i3 << 16) | i2;
        };
        refresh_int_thresholds();
        int sphase = 0;   // which survivor counter is in use
        // Exact evaluation of the queued survivors, 256 at a time; folds the candidate queues and refreshes the
        // thresholds.  Called by all threads (CTA-uniform), right after a barrier.
        auto drain = [&]() {
            const int ns = ctrl->nsurv[sphase];
            for (int base = 0; base < ns; base += kT) {
                const int s = base + tid;
                uint32_t idx = 0u, bits = 0u;
                float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
                if (s < ns) {
                    const uint32_t e = surv[s];
                    idx = e >> 4;
                    bits = e & 15u;
                    const QuadCode<M> cc = quad_load_code<M>(lp, idx, n);
#pragma unroll
                    for (int h = 0; h < M / 16; h++) {
                        const uint32_t cw[4] = {cc.v[h].x, cc.v[h].y, cc.v[h].z, cc.v[h].w};
#pragma unroll
                        for (int mm = 0; mm < 16; mm++) {
                            const int m = 16 * h + mm;
                            const uint32_t c = (cw[mm >> 2] >> (8 * (mm & 3))) & 255u;
                            const float4 t = lutf[m * 256 + c];     // plain (coherent) load: written by this CTA
                            a0 = __fadd_rn(a0, t.x);
                            a1 = __fadd_rn(a1, t.y);
                            a2 = __fadd_rn(a2, t.z);
                            a3 = __fadd_rn(a3, t.w);
                        }
                    }
                }
                const uint32_t b0 = __float_as_uint(a0), b1 = __float_as_uint(a1), b2 = __float_as_uint(a2),
                               b3 = __float_as_uint(a3);
                tk[0].push((bits & 1u) && b0 <= th0, make_key(b0, idx));
                tk[1].push((bits & 2u) && b1 <= th1, make_key(b1, idx));
                tk[2].push((bits & 4u) && b2 <= th2, make_key(b2, idx));
                tk[3].push((bits & 8u) && b3 <= th3, make_key(b3, idx));
                // at most kT new entries per queue and round: fold when another round could overflow
                const bool over = tk[0].pending() > kCap - kT || tk[1].pending() > kCap - kT ||
                                  tk[2].pending() > kCap - kT || tk[3].pending() > kCap - kT;
                if (__syncthreads_or(over) && base + kT < ns) {
#pragma unroll
                    for (int q = 0; q < 4; q++) tk[q].template flush<kT>(ext[q]);
                }
            }
            // every thread has read its survivors: this counter is next used after the NEXT drain, i.e. several
            // barriers from now, so it can be zeroed without one of its own
            if (tid == 0) ctrl->nsurv[sphase] = 0;
            sphase ^= 1;
            if (!topk_fold_small<kT, 4>(tk, ext)) {
#pragma unroll
                for (int q = 0; q < 4; q++) tk[q].template flush<kT>(ext[q]);
            }
            th0 = tk[0].threshold();
            th1 = tk[1].threshold();
            th2 = tk[2].threshold();
            th3 = tk[3].threshold();
            refresh_int_thresholds();
        };

        // a4: the filter.  Block b = codes b*kT + tid; every lane walks its own code (no carry between blocks).
        const uint32_t nblk = (n + kT - 1) / kT;
        const uint32_t vm01 = ((vmask & 1u) ? 0x8000u : 0u) | ((vmask & 2u) ? 0x80000000u : 0u);
        const uint32_t vm23 = ((vmask & 4u) ? 0x8000u : 0u) | ((vmask & 8u) ? 0x80000000u : 0u);
        QuadCode<M> c0 = quad_load_code<M>(lp, tid, n), c1 = quad_load_code<M>(lp, kT + tid, n), c2, c3;
#define QUAD_ITER(CUR, LOADTO, TB)                                                                   \
    {                                                                                                \
        LOADTO = quad_load_code<M>(lp, base + (TB + 2) * kT, n);                                     \
        uint2 lb = quad_block16(lutb, CUR.v[0], ws2, ws1, bs, loff);                                 \
        if constexpr (M == 32) {                                                                     \
            const uint2 hi = quad_block16(lutb + kQuadLutBytes, CUR.v[M / 16 - 1], ws2, ws1, bs, loff); \
            lb.x += hi.x;                                                                            \
            lb.y += hi.y;                                                                            \
        }                                                                                            \
        const uint32_t idx = base + TB * kT;                                                         \
        /* per half: 0x8000 + t - s keeps bit 15 iff s <= t (s, t < 2^15) */                         \
        const uint32_t m01 = (t01 - lb.x) & vm01, m23 = (t23 - lb.y) & vm23;                         \
        const bool hit = idx < n && (m01 | m23) != 0u;                                               \
        const unsigned bal = __ballot_sync(0xffffffffu, hit);                                        \
        if (bal) {                                                                                   \
            int slot = 0;                                                                            \
            const int leader = __ffs(bal) - 1;                                                       \
            if (lane == leader) slot = atomicAdd(&ctrl->nsurv[sphase], __popc(bal));                 \
            slot = __shfl_sync(0xffffffffu, slot, leader) + __popc(bal & lanemask_lt());             \
            if (hit)                                                                                 \
                surv[slot] = (idx << 4) | ((m01 >> 15) & 1u) | ((m01 >> 30) & 2u) | ((m23 >> 13) & 4u) | \
                             ((m23 >> 28) & 8u);                                                     \
        }                                                                                            \
    }
        // drain once enough survivors are waiting (tight thresholds early are worth more than fewer drains); checked
        // every 1024 codes (kTileBlocks blocks), except after the last block, which goes straight to the final drain
#define QUAD_CHECK(NEXT)                                                                             \
    if ((NEXT) % Cfg::kTileBlocks == 0 && t0 + (NEXT) < nblk) {                                      \
        const int seen = *reinterpret_cast<volatile int*>(&ctrl->nsurv[sphase]);                     \
        if (__syncthreads_or(seen > p.quad_drain_at)) drain();                                       \
    }
        for (uint32_t t0 = 0; t0 < nblk; t0 += 4) {
            const uint32_t base = t0 * kT + tid;
            QUAD_ITER(c0, c2, 0)
            QUAD_CHECK(1)
            if (t0 + 1 < nblk) QUAD_ITER(c1, c3, 1)
            QUAD_CHECK(2)
            if (t0 + 2 < nblk) QUAD_ITER(c2, c0, 2)
            QUAD_CHECK(3)
            if (t0 + 3 < nblk) QUAD_ITER(c3, c1, 3)
            QUAD_CHECK(4)
        }
#undef QUAD_CHECK
#undef QUAD_ITER
        __syncthreads();
        drain();
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (vmask & (1u << q)) {
                const int nb = tk[q].count();
                const uint64_t* s = tk[q].sorted();
                for (int i = tid; i < nb; i += kT) p.out_keys[static_cast<int64_t>(pair[q]) * p.k + i] = s[i];
                if (tid == 0) {
                    p.out_cnt[pair[q]] = nb;
                    if (nb == p.k) atomicMin(p.qthr + qi[q], static_cast<uint32_t>(s[p.k - 1] >> 32));
                }
            }
        }
        // no barrier here: the one at the top of the next item separates these reads from its first writes
    }
}

template <int M, int DSUB>
int launch_scan_quad_t(const ScanParams& sp, const float* pq_t, int64_t grid, cudaStream_t st) {
    size_t smem = quad_smem_bytes(M, sp.d, sp.k);
    auto kernel = scan_quad_kernel<M, DSUB>;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    kernel<<<(unsigned)grid, QuadCfg<M>::kT, smem, st>>>(sp, pq_t);
    return cudaPeekAtLastError() == cudaSuccess ? 0 : -1;
}

// CTAs the kernel will be launched with (the caller sizes the exact-LUT scratch with it): 0 when it does not fit
template <int M, int DSUB>
int quad_grid_t(int d, int k, int64_t npairs, int num_sms) {
    size_t smem = quad_smem_bytes(M, d, k);
    auto kernel = scan_quad_kernel<M, DSUB>;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, QuadCfg<M>::kT, smem) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        return 0;
    }
    int64_t grid = static_cast<int64_t>(per_sm) * num_sms;
    if (grid > npairs) grid = npairs;
    return static_cast<int>(grid);
}

inline int quad_grid(int M, int dsub, int d, int k, int64_t npairs, int num_sms) {
    if (quad_smem_bytes(M, d, k) > 227 * 1024) return 0;
    if (M == 32) {
        switch (dsub) {
            case 4: return quad_grid_t<32, 4>(d, k, npairs, num_sms);
            case 8: return quad_grid_t<32, 8>(d, k, npairs, num_sms);
            default: return quad_grid_t<32, 0>(d, k, npairs, num_sms);
        }
    }
    switch (dsub) {
        case 4: return quad_grid_t<16, 4>(d, k, npairs, num_sms);
        case 6: return quad_grid_t<16, 6>(d, k, npairs, num_sms);
        case 8: return quad_grid_t<16, 8>(d, k, npairs, num_sms);
        default: return quad_grid_t<16, 0>(d, k, npairs, num_sms);
    }
}

// returns 0, or -1 on a launch error (caller reads cudaGetLastError)
inline int launch_scan_quad(const ScanParams& sp, const float* pq_t, int64_t grid, cudaStream_t st) {
    if (sp.M == 32) {
        switch (sp.dsub) {
            case 4: return launch_scan_quad_t<32, 4>(sp, pq_t, grid, st);
            case 8: return launch_scan_quad_t<32, 8>(sp, pq_t, grid, st);
            default: return launch_scan_quad_t<32, 0>(sp, pq_t, grid, st);
        }
    }
    switch (sp.dsub) {
        case 4: return launch_scan_quad_t<16, 4>(sp, pq_t, grid, st);
        case 6: return launch_scan_quad_t<16, 6>(sp, pq_t, grid, st);
        case 8: return launch_scan_quad_t<16, 8>(sp, pq_t, grid, st);
        default: return launch_scan_quad_t<16, 0>(sp, pq_t, grid, st);
    }
}

}  // namespace b200
