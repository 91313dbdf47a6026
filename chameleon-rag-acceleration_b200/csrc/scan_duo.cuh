// scan_duo.cuh -- the skewed-lane ADC scan (scan_skew.cuh) for TWO queries that probe the SAME list (M = 16).
//
// At batch sizes where every list is probed by several queries (10k-query batches: ~39 queries per list on
// IVF8192, nprobe 32) the scan is bound by instruction issue and the shared-memory pipe, not by HBM.  This kernel
// halves the per-code overhead by pairing queries of a list:
//   * the LUTs of the two queries are interleaved entry-wise, lut2[c][e] = (T_a[e % 16][c], T_b[e % 16][c]),
//     e = 0..31, row stride 256 B: one PRMT builds the address (c << 8) | 8*(lane % 16), ONE LDS.64 returns both
//     table entries already packed as an f32x2 operand.  A 64-bit shared load is served per half-warp; the 16 lanes
//     of a half-warp read entries (lane % 16) + p -- 16 distinct bank pairs for ANY code bytes -- so the gather stays
//     conflict-free and the two LUTs together still take 64 KB (two CTAs per SM);
//   * code loads, the byte-rotated window and the address PRMT are shared by the two queries;
//   * the lane-dependent "restart at m = 0 / capture at m = 15" of the skewed scan is expressed with two packed
//     accumulators: F collects the code that finishes in this 16-step block, N the code that starts in it; each
//     step is ISETP + two predicated FADD2 (add.rn.f32x2 rounds each half exactly like the scalar add).  At the
//     end of the block F holds both queries' finished distances and N becomes the next block's F.  0 + T is exact,
//     and every lane still adds its code's 16 entries in ascending m: distances stay bit-identical to the oracle.
// Per 16-step block and 2 x 512 lookups: 16 PRMT + 16 LDS.64 + 16 ISETP + 32 FADD2 (+ window, loads, pushes)
// instead of 2 x (16 PRMT + 16 LDS + 32 FFMA).
//
// Work items are GROUPS (pair_a, pair_b) of (query, probe) pairs with the same list, produced by the pair setup
// (kernels.cuh pair_scatter_kernel); pair_b = -1 for the odd one out.
//
// Reference semantics: ADC.hpp:75-99 / IVFPQ_1B_search.ipynb:7948-7960 (sum over m ascending),
// LUT_construction.hpp:180-209 / ipynb:7929-7946 (LUT), priority_queue_L1.hpp:65-75 (strict <).
#pragma once
#include "scan_skew.cuh"

namespace b200 {

constexpr int kDuoRowEntries = 32;                       // (T_a, T_b) pairs per code value, period 16
constexpr int kDuoLutBytes = 256 * kDuoRowEntries * 8;   // 64 KB for both queries
constexpr int kDuoTB = 4;                                // blocks (of 32 codes) per warp per tile
constexpr int kDuoCap = 2 * kThreads * kDuoTB;           // candidate queue per query: two tiles

inline bool duo_supported(int M, int d, int k) {
    (void)d;
    return M == 16 && k <= B200_IVFPQ_MAX_K;
}

__host__ __device__ inline size_t duo_smem_bytes(int d, int k) {
    return kDuoLutBytes + 2 * sizeof(float) * static_cast<size_t>((d + 3) & ~3) + 2 * TopK::smem_bytes(k, kDuoCap) + 16 +
           2 * sizeof(DuoGroup);
}

__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

template <int B>
__device__ __forceinline__ uint64_t duo_lookup(const char* __restrict__ lutb, uint32_t w, uint32_t loff, int p) {
    // byte0 = loff (8 * (lane % 16) < 128), byte1 = byte B of w, bytes 2,3 = 0  ->  (c << 8) | 8*(lane % 16)
    uint32_t a = __byte_perm(w, loff, 0x6504 | (B << 4));
    return *reinterpret_cast<const uint64_t*>(lutb + a + 8 * p);
}

// One block = 16 lookups per lane and query over the byte window [r, r+16) of cur|nxt.  acc holds the partial sums
// (a, b) of the code the lane is on; returns the finished distances of the code that completed in this block.
// keep2[p] = (0, 0) at the step where the lane starts a new code (m == 0), else (1, 1); cap2[p] = (1, 1) at the step
// where it finishes one (m == 15), else (0, 0):  acc = acc*keep + T, fin = acc*cap + fin are exact (x*1 + y and
// x*0 + y round like y + x and y), two FFMA2 per step for both queries.
__device__ __forceinline__ uint64_t duo_block16(const char* __restrict__ lutb, const uint4& cur, const uint4& nxt,
                                                bool ws2, bool ws1, uint32_t bs, uint32_t loff,
                                                const uint64_t (&keep2)[16], const uint64_t (&cap2)[16],
                                                uint64_t& acc) {
    const uint32_t y0 = ws2 ? cur.z : cur.x, y1 = ws2 ? cur.w : cur.y, y2 = ws2 ? nxt.x : cur.z,
                   y3 = ws2 ? nxt.y : cur.w, y4 = ws2 ? nxt.z : nxt.x, y5 = ws2 ? nxt.w : nxt.y;
    const uint32_t z0 = ws1 ? y1 : y0, z1 = ws1 ? y2 : y1, z2 = ws1 ? y3 : y2, z3 = ws1 ? y4 : y3,
                   z4 = ws1 ? y5 : y4;
    const uint32_t w0 = __funnelshift_r(z0, z1, bs), w1 = __funnelshift_r(z1, z2, bs),
                   w2 = __funnelshift_r(z2, z3, bs), w3 = __funnelshift_r(z3, z4, bs);
    uint64_t fin = 0ull;
#define DUO_STEP(W, B, P)                                   \
    {                                                       \
        const uint64_t T = duo_lookup<B>(lutb, W, loff, P); \
        acc = fma_f32x2(acc, keep2[P], T);                  \
        fin = fma_f32x2(acc, cap2[P], fin);                 \
    }
    DUO_STEP(w0, 0, 0) DUO_STEP(w0, 1, 1) DUO_STEP(w0, 2, 2) DUO_STEP(w0, 3, 3)
    DUO_STEP(w1, 0, 4) DUO_STEP(w1, 1, 5) DUO_STEP(w1, 2, 6) DUO_STEP(w1, 3, 7)
    DUO_STEP(w2, 0, 8) DUO_STEP(w2, 1, 9) DUO_STEP(w2, 2, 10) DUO_STEP(w2, 3, 11)
    DUO_STEP(w3, 0, 12) DUO_STEP(w3, 1, 13) DUO_STEP(w3, 2, 14) DUO_STEP(w3, 3, 15)
#undef DUO_STEP
    return fin;
}

// 32-byte descriptor, global -> shared, asynchronously (LDGSTS); completion: cp.async.wait_all + barrier
__device__ __forceinline__ void duo_copy_group_async(DuoGroup* dst, const DuoGroup* src) {
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d + 16), "l"(reinterpret_cast<const char*>(src) + 16)
                 : "memory");
}

template <int DSUB>
__device__ __forceinline__ float lut_entry_regs(const float (&pv)[DSUB], const uint64_t (&rr2)[DSUB / 2]) {
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

// NOTE on packed arithmetic: ptxas contracts mul.rn.f32x2 followed by add.rn.f32x2 into FFMA2 (even with
// -fmad=false), which would break bit-parity with the oracle's separately rounded multiply and add (the parity
// tests compare every distance bit for bit and caught exactly this).  The square is therefore written as
// fma(d, d, -0.0) with the -0.0 coming from a kernel parameter (opaque to ptxas): x*y + (-0.0) rounds exactly like
// x*y, and an FFMA2 cannot be fused with the add that follows.
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// (T_a, T_b) for one (m, c) entry, packed across the two QUERIES: rab[j] = (r_a[j], r_b[j]); the PQ centroid value
// p_j is broadcast to both halves (FADD2's scalar operand form).  Per dimension: FADD2 (sub), FFMA2 (square), FADD2
// (accumulate) -- each half rounds exactly like the oracle's scalar sub / mul / add, j ascending.
template <int DSUB>
__device__ __forceinline__ uint64_t lut_entry_duo(const float (&pv)[DSUB], const uint64_t (&rab)[DSUB],
                                                  uint64_t negzero2) {
    uint64_t acc = 0ull;
#pragma unroll
    for (int j = 0; j < DSUB; j++) {
        const uint64_t d2 = sub_f32x2(rab[j], pack_f32x2(pv[j], pv[j]));
        acc = add_f32x2(acc, fma_f32x2(d2, d2, negzero2));
    }
    return acc;
}

// DSUB = d / 16 when it is one of the specialised values (residual slices held in registers), 0 = generic.
template <int DSUB>
__global__ void __launch_bounds__(kThreads, 2) scan_duo16_kernel(const ScanParams p, const float* __restrict__ pq_t) {
    constexpr int M = 16;
    extern __shared__ __align__(1024) unsigned char smem_duo[];
    uint64_t* lut2 = reinterpret_cast<uint64_t*>(smem_duo);
    const int dpad = (p.d + 3) & ~3;
    float* res_a = reinterpret_cast<float*>(smem_duo + kDuoLutBytes);
    float* res_b = res_a + dpad;
    TopK tka, tkb;
    tka.bind(res_b + dpad, p.k, kDuoCap);
    tkb.bind(reinterpret_cast<unsigned char*>(res_b + dpad) + TopK::smem_bytes(p.k, kDuoCap), p.k, kDuoCap);
    int* s_work = tkb.meta + 4;
    const char* lutb = reinterpret_cast<const char*>(lut2);

    const int tid = threadIdx.x, lane = tid & 31;
    const int r = lane & (M - 1);
    const int pstart = (M - r) & (M - 1);   // step at which this lane starts a new code (m == 0)
    const int pend = M - 1 - r;             // step at which this lane finishes a code (m == 15)
    uint64_t keep2[M], cap2[M];
#pragma unroll
    for (int s = 0; s < M; s++) {
        keep2[s] = (s == pstart) ? 0ull : 0x3f8000003f800000ull;   // (1.0f, 1.0f)
        cap2[s] = (s == pend) ? 0x3f8000003f800000ull : 0ull;
    }
    const uint32_t loff = static_cast<uint32_t>(r) * 8u;
    const bool ws2 = (r & 8) != 0, ws1 = (r & 4) != 0;
    const uint32_t bs = static_cast<uint32_t>(r & 3) * 8u;
    const int ngroups = p.stats->ngroups;
    const int dsub = DSUB ? DSUB : p.dsub;
    // LUT build mapping: this thread owns sub-quantizer lm and code values lc0 + 16 i; a half-warp (same code value,
    // lm = 0..15) stores 16 consecutive 8-byte entries: conflict-free
    const int lm = tid & (M - 1), lc0 = tid >> 4;

    // Work items are pulled from a global counter, two ahead: the counter value for item i+1 is fetched at the top of
    // item i, and once it has arrived (after the LUT build) thread 0 starts an asynchronous copy of that item's
    // descriptor into shared memory, so that the next iteration starts without a dependent global load.
    DuoGroup* s_grp = reinterpret_cast<DuoGroup*>(s_work + 4);   // [2], 16-byte aligned
    int next_work = 0, buf = 0;
    if (tid == 0) {
        next_work = atomicAdd(&p.stats->work_counter, 1);
        if (next_work < ngroups) duo_copy_group_async(&s_grp[0], static_cast<const DuoGroup*>(p.groups) + next_work);
    }
    for (;;) {
        if (tid == 0) {
            *s_work = next_work;
            asm volatile("cp.async.wait_all;" ::: "memory");
        }
        __syncthreads();
        const int wk = *s_work;
        if (wk >= ngroups) break;
        if (tid == 0) next_work = atomicAdd(&p.stats->work_counter, 1);
        const DuoGroup grp = s_grp[buf];
        const int pair_a = grp.pair_a;
        const bool has_b = grp.pair_b >= 0;
        const int pair_b = has_b ? grp.pair_b : grp.pair_a;
        const int qa = pair_a / p.nprobe, qb = pair_b / p.nprobe;
        const int list = grp.list;
        const int64_t beg = grp.beg;
        const uint32_t n = grp.n;                                 // > 0: empty lists form no groups
        const uint32_t n_b = has_b ? n : 0u;
        const uint4* lp = reinterpret_cast<const uint4*>(p.codes + beg * M);

        // lane's code in block b is b*256 + tid.  Ring of four code registers: slot (b+1)%4 holds code b.
        uint4 c0 = make_uint4(0u, 0u, 0u, 0u);                    // "code -1" of the prologue block
        uint4 c1 = skew_load_code16(lp, tid, n);                  // code 0
        uint4 c2 = skew_load_code16(lp, 256u + tid, n);           // code 1
        uint4 c3;

        // LUT build, specialised path: the PQ centroid slices of entries i (code value lc0 + 16 i) are loaded two
        // entries ahead of their use; the first two do not even wait for the residuals
#define DUO_LUT_LOAD(PV, I)                                                                        \
    {                                                                                              \
        const float* pc_ = pq_t + static_cast<int64_t>(lc0 + 16 * (I)) * (DSUB * M) + lm;          \
        _Pragma("unroll") for (int j = 0; j < DSUB; j++) PV[j] = __ldg(pc_ + j * M);               \
    }
#define DUO_LUT_EMIT(PV, I)                                                                        \
    {                                                                                              \
        const uint64_t e_ = lut_entry_duo<DSUB>(PV, rab, negzero2);                                \
        uint64_t* row_ = lut2 + (lc0 + 16 * (I)) * kDuoRowEntries + lm;                            \
        row_[0] = e_;                                                                              \
        row_[16] = e_;                                                                             \
    }
        float pv0[DSUB ? DSUB : 1], pv1[DSUB ? DSUB : 1], pv2[DSUB ? DSUB : 1], pv3[DSUB ? DSUB : 1];
        if constexpr (DSUB != 0) {
            DUO_LUT_LOAD(pv0, 0)
            DUO_LUT_LOAD(pv1, 1)
        }
        // a2: residuals of both queries against the list's centroid.  Specialised path: stored as pairs (r_a, r_b),
        // j-major (pair j * 16 + m), so that a half-warp reads 16 consecutive pairs in the LUT build
        for (int j = tid; j < p.d; j += kThreads) {
            const float cj = p.cent[static_cast<int64_t>(list) * p.d + j];
            const float ra = __fsub_rn(p.xq[static_cast<int64_t>(qa) * p.d + j], cj);
            const float rb = __fsub_rn(p.xq[static_cast<int64_t>(qb) * p.d + j], cj);
            if (DSUB) {
                reinterpret_cast<uint64_t*>(res_a)[(j % (DSUB ? DSUB : 1)) * M + j / (DSUB ? DSUB : 1)] = pack_f32x2(ra, rb);
            } else {
                res_a[j] = ra;
                res_b[j] = rb;
            }
        }
        const uint32_t ext_a = *reinterpret_cast<volatile uint32_t*>(p.qthr + qa);
        const uint32_t ext_b = *reinterpret_cast<volatile uint32_t*>(p.qthr + qb);
        if (tid == 0) {
            tka.reset(ext_a);
            tkb.reset(ext_b);
        }
        __syncthreads();
        // a3: both LUTs, interleaved; the PQ centroid slice is loaded once for the two queries
        if constexpr (DSUB != 0) {
            uint64_t rab[DSUB];
#pragma unroll
            for (int j = 0; j < DSUB; j++) rab[j] = reinterpret_cast<const uint64_t*>(res_a)[j * M + lm];
            const uint64_t negzero2 = p.negzero2;
#pragma unroll 1
            for (int i = 0; i < 16; i += 4) {
                DUO_LUT_LOAD(pv2, i + 2)
                DUO_LUT_LOAD(pv3, i + 3)
                DUO_LUT_EMIT(pv0, i)
                DUO_LUT_EMIT(pv1, i + 1)
                if (i + 4 < 16) {
                    DUO_LUT_LOAD(pv0, i + 4)
                    DUO_LUT_LOAD(pv1, i + 5)
                }
                DUO_LUT_EMIT(pv2, i + 2)
                DUO_LUT_EMIT(pv3, i + 3)
            }
#undef DUO_LUT_LOAD
#undef DUO_LUT_EMIT
        } else {
            for (int i = 0; i < 16; i++) {
                const int c = lc0 + 16 * i;
                const float* pc = pq_t + static_cast<int64_t>(c) * dsub * M + lm;
                const float* ra = res_a + lm * dsub;
                const float* rb = res_b + lm * dsub;
                float a = 0.0f, b = 0.0f;
                for (int j = 0; j < dsub; j++) {
                    const float pj = __ldg(pc + j * M);
                    a = sqdiff_acc(a, ra[j], pj);
                    b = sqdiff_acc(b, rb[j], pj);
                }
                const uint64_t e = pack_f32x2(a, b);
                uint64_t* row = lut2 + c * kDuoRowEntries + lm;
                row[0] = e;
                row[16] = e;
            }
        }
        __syncthreads();
        buf ^= 1;
        if (tid == 0 && next_work < ngroups) duo_copy_group_async(&s_grp[buf], static_cast<const DuoGroup*>(p.groups) + next_work);

        // a4 + a5.  Iteration `it` processes block b = it - 1 (it = 0 is the prologue that only feeds bytes
        // 0..r-1 of code 0).  Four iterations per tile, fully unrolled so that the code ring needs no moves.
        uint32_t thr_a = ext_a, thr_b = ext_b;
        uint64_t carry = 0ull;
        const uint32_t nblk = (n + 255u) >> 8;
#define DUO_ITER(CUR, NXT, LOADTO, TB)                                                          \
    {                                                                                           \
        LOADTO = skew_load_code16(lp, base + (TB + 2) * 256u, n);                               \
        const uint64_t fin = duo_block16(lutb, CUR, NXT, ws2, ws1, bs, loff, keep2, cap2, carry); \
        const uint32_t idx = base + TB * 256u - 256u; /* wraps past 2^32 for the prologue block */ \
        const uint32_t ba = static_cast<uint32_t>(fin), bb = static_cast<uint32_t>(fin >> 32);  \
        const bool pa = idx < n && ba <= thr_a, pb = idx < n_b && bb <= thr_b;                  \
        if (__any_sync(0xffffffffu, pa || pb)) { /* rare once a threshold is known */           \
            tka.push(pa, make_key(ba, idx));                                                    \
            tkb.push(pb, make_key(bb, idx));                                                    \
        }                                                                                       \
    }
        uint32_t t0 = 0;
        for (; t0 + (kDuoTB - 1) <= nblk; t0 += kDuoTB) {   // full tiles: iterations t0 .. t0 + 3
            const uint32_t base = t0 * 256u + tid;          // code index of block b = t0 (iteration t0 + 1)
            DUO_ITER(c0, c1, c3, 0)
            DUO_ITER(c1, c2, c0, 1)
            DUO_ITER(c2, c3, c1, 2)
            DUO_ITER(c3, c0, c2, 3)
            // one barrier serves both queues; no threshold yet (cold start) -> fold the first tile in right away
            const int lim_a = thr_a == kInfBits ? 0 : kDuoCap - kThreads * kDuoTB;
            const int lim_b = thr_b == kInfBits ? 0 : kDuoCap - kThreads * kDuoTB;
            const int seen_a = *reinterpret_cast<volatile int*>(&tka.meta[1]);
            const int seen_b = *reinterpret_cast<volatile int*>(&tkb.meta[1]);
            if (__syncthreads_or(seen_a > lim_a || seen_b > lim_b)) {
                tka.flush<kThreads>(ext_a);
                tkb.flush<kThreads>(ext_b);
            }
            thr_a = tka.threshold();
            thr_b = tkb.threshold();
        }
        if (t0 <= nblk) {   // remaining 1..3 iterations (CTA-uniform); the queues have room for a whole tile
            const uint32_t base = t0 * 256u + tid;
            DUO_ITER(c0, c1, c3, 0)
            if (t0 + 1 <= nblk) DUO_ITER(c1, c2, c0, 1)
            if (t0 + 2 <= nblk) DUO_ITER(c2, c3, c1, 2)
        }
#undef DUO_ITER
        __syncthreads();
        tka.flush<kThreads>(ext_a);
        tkb.flush<kThreads>(ext_b);
        {
            const int nb = tka.count();
            const uint64_t* s = tka.sorted();
            for (int i = tid; i < nb; i += kThreads) p.out_keys[static_cast<int64_t>(pair_a) * p.k + i] = s[i];
            if (tid == 0) {
                p.out_cnt[pair_a] = nb;
                if (nb == p.k) atomicMin(p.qthr + qa, static_cast<uint32_t>(s[p.k - 1] >> 32));
            }
        }
        if (has_b) {
            const int nb = tkb.count();
            const uint64_t* s = tkb.sorted();
            for (int i = tid; i < nb; i += kThreads) p.out_keys[static_cast<int64_t>(pair_b) * p.k + i] = s[i];
            if (tid == 0) {
                p.out_cnt[pair_b] = nb;
                if (nb == p.k) atomicMin(p.qthr + qb, static_cast<uint32_t>(s[p.k - 1] >> 32));
            }
        }
        // no barrier here: the one at the top of the next item separates these reads from its first writes
    }
}

template <int DSUB>
int launch_scan_duo_t(const ScanParams& sp, const float* pq_t, int64_t npairs, int num_sms, cudaStream_t st) {
    size_t smem = duo_smem_bytes(sp.d, sp.k);
    auto kernel = scan_duo16_kernel<DSUB>;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem) != cudaSuccess) return -1;
    if (per_sm < 1) return -1;
    int64_t grid = static_cast<int64_t>(per_sm) * num_sms;
    if (grid > npairs) grid = npairs;
    kernel<<<(unsigned)grid, kThreads, smem, st>>>(sp, pq_t);
    return cudaPeekAtLastError() == cudaSuccess ? 0 : -1;
}

// returns 0, or -1 on a launch error (caller reads cudaGetLastError)
inline int launch_scan_duo(const ScanParams& sp, const float* pq_t, int64_t npairs, int num_sms, cudaStream_t st) {
    switch (sp.dsub) {
        case 4: return launch_scan_duo_t<4>(sp, pq_t, npairs, num_sms, st);
        case 6: return launch_scan_duo_t<6>(sp, pq_t, npairs, num_sms, st);
        case 8: return launch_scan_duo_t<8>(sp, pq_t, npairs, num_sms, st);
        case 16: return launch_scan_duo_t<16>(sp, pq_t, npairs, num_sms, st);
        default: return launch_scan_duo_t<0>(sp, pq_t, npairs, num_sms, st);
    }
}

}  // namespace b200
