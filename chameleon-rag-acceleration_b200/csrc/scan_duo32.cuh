// scan_duo32.cuh -- the two-query skewed-lane ADC scan (scan_duo.cuh) for M = 32 (C5: IVF8192,PQ32; the reference's
// RALM-S / SIFT M32 accelerator shapes).
//
// A 32-byte code is walked as two 16-byte halves with the 16-phase skew of the M = 16 kernel: lane phase r = lane % 16,
// step p of a 16-step block looks at entry e = r + p of a 32-entry LUT row (row = code value, 256 B, (T_a, T_b) pairs
// read with one conflict-free LDS.64).  Which sub-quantizer entry e stands for alternates with the block:
//   even block  window = bytes [r, r+16) of  lo(code i) | hi(code i)      entry e -> m = e            (0..30, code i)
//   odd  block  window = bytes [r, r+16) of  hi(code i) | lo(code i+1)    entry e -> m = (e + 16) % 32
//                                                                          (e < 16: m = 16 + e of code i,
//                                                                           e >= 16: m = e - 16 of code i + 1)
// so there are two tables, lutE[c][e] = T[e][c] and lutO[c][e] = T[(e + 16) % 32][c] -- no periodic copies at all,
// 2 x 64 KB for the two queries together (one CTA of 512 threads per SM).  Every lane adds its code's 32 entries in
// ascending m: bit-identical to the oracle.  Only the odd block has the lane-dependent restart (m = 0) and capture
// (m = 31), as exact FFMA2 with 0/1 multipliers; the even block is 15 plain FADD2 plus one FFMA2 that restarts the
// lanes with r = 0.  Per look-up pair: PRMT + LDS.64 + 1.5 FFMA2/FADD2.
//
// Reference semantics: ADC.hpp:75-99 (M = 32 accelerator) / IVFPQ_1B_search.ipynb:7948-7960 (sum over m ascending),
// LUT_construction.hpp:180-209 / ipynb:7929-7946 (LUT), priority_queue_L1.hpp:65-75 (strict <).
#pragma once
#include "scan_duo.cuh"

namespace b200 {

constexpr int kDuo32Threads = 512;
constexpr int kDuo32TableBytes = 256 * 32 * 8;             // one table: 256 code values x 32 entries x (T_a, T_b)
constexpr int kDuo32Unroll = 3;                            // codes per lane per tile (ring of three code registers)
constexpr int kDuo32Tile = kDuo32Threads * kDuo32Unroll;   // codes per tile
constexpr int kDuo32Cap = 2048;                            // candidate queue per query

inline bool duo32_supported(int M, int d, int k) {
    (void)d;
    return M == 32 && k <= B200_IVFPQ_MAX_K;
}

__host__ __device__ inline size_t duo32_smem_bytes(int d, int k) {
    return 2 * static_cast<size_t>(kDuo32TableBytes) + 2 * sizeof(float) * static_cast<size_t>((d + 3) & ~3) +
           2 * TopK::smem_bytes(k, kDuo32Cap) + 16 + 2 * sizeof(DuoGroup);
}

struct Code32 {
    uint4 lo, hi;
};

__device__ __forceinline__ Code32 duo32_load_code(const uint4* __restrict__ lp, uint32_t idx, uint32_t n) {
    Code32 c;
    c.lo = make_uint4(0u, 0u, 0u, 0u);
    c.hi = c.lo;
    if (idx < n) {
        c.lo = __ldg(lp + 2 * static_cast<size_t>(idx));
        c.hi = __ldg(lp + 2 * static_cast<size_t>(idx) + 1);
    }
    return c;
}

// byte window [r, r+16) of cur|nxt as four words (same construction as the M = 16 kernel)
__device__ __forceinline__ void duo32_window(const uint4& cur, const uint4& nxt, bool ws2, bool ws1, uint32_t bs,
                                             uint32_t (&w)[4]) {
    const uint32_t y0 = ws2 ? cur.z : cur.x, y1 = ws2 ? cur.w : cur.y, y2 = ws2 ? nxt.x : cur.z,
                   y3 = ws2 ? nxt.y : cur.w, y4 = ws2 ? nxt.z : nxt.x, y5 = ws2 ? nxt.w : nxt.y;
    const uint32_t z0 = ws1 ? y1 : y0, z1 = ws1 ? y2 : y1, z2 = ws1 ? y3 : y2, z3 = ws1 ? y4 : y3,
                   z4 = ws1 ? y5 : y4;
    w[0] = __funnelshift_r(z0, z1, bs);
    w[1] = __funnelshift_r(z1, z2, bs);
    w[2] = __funnelshift_r(z2, z3, bs);
    w[3] = __funnelshift_r(z3, z4, bs);
}

// even block: all 16 entries belong to the lane's current code; lanes with r == 0 start it here (keep0 = (0, 0))
__device__ __forceinline__ void duo32_even(const char* __restrict__ lut_e, const uint4& lo, const uint4& hi, bool ws2,
                                           bool ws1, uint32_t bs, uint32_t loff, uint64_t keep0, uint64_t& acc) {
    uint32_t w[4];
    duo32_window(lo, hi, ws2, ws1, bs, w);
    acc = fma_f32x2(acc, keep0, duo_lookup<0>(lut_e, w[0], loff, 0));
#define DUO32_ADD(W, B, P) acc = add_f32x2(acc, duo_lookup<B>(lut_e, W, loff, P));
    DUO32_ADD(w[0], 1, 1) DUO32_ADD(w[0], 2, 2) DUO32_ADD(w[0], 3, 3)
    DUO32_ADD(w[1], 0, 4) DUO32_ADD(w[1], 1, 5) DUO32_ADD(w[1], 2, 6) DUO32_ADD(w[1], 3, 7)
    DUO32_ADD(w[2], 0, 8) DUO32_ADD(w[2], 1, 9) DUO32_ADD(w[2], 2, 10) DUO32_ADD(w[2], 3, 11)
    DUO32_ADD(w[3], 0, 12) DUO32_ADD(w[3], 1, 13) DUO32_ADD(w[3], 2, 14) DUO32_ADD(w[3], 3, 15)
#undef DUO32_ADD
}

// odd block: finishes code i (capture at m = 31) and starts code i + 1 (restart at m = 0); returns the finished
// distances (a, b) of code i
__device__ __forceinline__ uint64_t duo32_odd(const char* __restrict__ lut_o, const uint4& hi, const uint4& nlo, bool ws2,
                                              bool ws1, uint32_t bs, uint32_t loff, const uint64_t (&keep2)[16],
                                              const uint64_t (&cap2)[16], uint64_t& acc) {
    uint32_t w[4];
    duo32_window(hi, nlo, ws2, ws1, bs, w);
    uint64_t fin = 0ull;
#define DUO32_STEP(W, B, P)                                   \
    {                                                         \
        const uint64_t T = duo_lookup<B>(lut_o, W, loff, P);  \
        acc = fma_f32x2(acc, keep2[P], T);                    \
        fin = fma_f32x2(acc, cap2[P], fin);                   \
    }
    DUO32_STEP(w[0], 0, 0) DUO32_STEP(w[0], 1, 1) DUO32_STEP(w[0], 2, 2) DUO32_STEP(w[0], 3, 3)
    DUO32_STEP(w[1], 0, 4) DUO32_STEP(w[1], 1, 5) DUO32_STEP(w[1], 2, 6) DUO32_STEP(w[1], 3, 7)
    DUO32_STEP(w[2], 0, 8) DUO32_STEP(w[2], 1, 9) DUO32_STEP(w[2], 2, 10) DUO32_STEP(w[2], 3, 11)
    DUO32_STEP(w[3], 0, 12) DUO32_STEP(w[3], 1, 13) DUO32_STEP(w[3], 2, 14) DUO32_STEP(w[3], 3, 15)
#undef DUO32_STEP
    return fin;
}

// DSUB = d / 32 when it is one of the specialised values (residual pairs held in registers), 0 = generic.
template <int DSUB>
__global__ void __launch_bounds__(kDuo32Threads, 1) scan_duo32_kernel(const ScanParams p, const float* __restrict__ pq_t) {
    constexpr int M = 32;
    extern __shared__ __align__(1024) unsigned char smem_duo32[];
    uint64_t* lut_e = reinterpret_cast<uint64_t*>(smem_duo32);
    uint64_t* lut_o = lut_e + 256 * 32;
    const int dpad = (p.d + 3) & ~3;
    uint64_t* res_ab = reinterpret_cast<uint64_t*>(smem_duo32 + 2 * kDuo32TableBytes);   // [d] pairs (r_a, r_b)
    TopK tka, tkb;
    unsigned char* tk_base = reinterpret_cast<unsigned char*>(res_ab + dpad);
    tka.bind(tk_base, p.k, kDuo32Cap);
    tkb.bind(tk_base + TopK::smem_bytes(p.k, kDuo32Cap), p.k, kDuo32Cap);
    int* s_work = tkb.meta + 4;
    DuoGroup* s_grp = reinterpret_cast<DuoGroup*>(s_work + 4);
    const char* lute_b = reinterpret_cast<const char*>(lut_e);
    const char* luto_b = reinterpret_cast<const char*>(lut_o);

    const int tid = threadIdx.x, lane = tid & 31;
    const int r = lane & 15;
    const int pstart = (16 - r) & 15;        // odd block: step at which the lane starts its next code (r > 0)
    const int pend = 15 - r;                 // odd block: step at which the lane finishes its code (m == 31)
    uint64_t keep2[16], cap2[16];
#pragma unroll
    for (int s = 0; s < 16; s++) {
        keep2[s] = (r != 0 && s == pstart) ? 0ull : 0x3f8000003f800000ull;
        cap2[s] = (s == pend) ? 0x3f8000003f800000ull : 0ull;
    }
    const uint64_t keep0 = r == 0 ? 0ull : 0x3f8000003f800000ull;   // even block, step 0: lanes with r == 0 restart
    const uint32_t loff = static_cast<uint32_t>(r) * 8u;
    const bool ws2 = (r & 8) != 0, ws1 = (r & 4) != 0;
    const uint32_t bs = static_cast<uint32_t>(r & 3) * 8u;
    const int ngroups = p.stats->ngroups;
    const int dsub = DSUB ? DSUB : p.dsub;
    // LUT build mapping: sub-quantizer lm = tid % 32, code values lc0 + 16 i: a warp stores 32 consecutive entries
    const int lm = tid & 31, lc0 = tid >> 5;

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
        const uint32_t n = grp.n;
        const uint32_t n_b = has_b ? n : 0u;
        const uint4* lp = reinterpret_cast<const uint4*>(p.codes + grp.beg * M);

        // lane's code in iteration it is it * 512 + tid; ring of three codes
        Code32 c0;                                                // "code -1" of the prologue
        c0.lo = make_uint4(0u, 0u, 0u, 0u);
        c0.hi = c0.lo;
        Code32 c1 = duo32_load_code(lp, tid, n);                  // code 0
        Code32 c2;

        // a2: residual pairs, j-major (pair j * 32 + m) on the specialised path
        for (int j = tid; j < p.d; j += kDuo32Threads) {
            const float cj = p.cent[static_cast<int64_t>(list) * p.d + j];
            const float ra = __fsub_rn(p.xq[static_cast<int64_t>(qa) * p.d + j], cj);
            const float rb = __fsub_rn(p.xq[static_cast<int64_t>(qb) * p.d + j], cj);
            const int dst = DSUB ? (j % (DSUB ? DSUB : 1)) * M + j / (DSUB ? DSUB : 1) : j;
            res_ab[dst] = pack_f32x2(ra, rb);
        }
        const uint32_t ext_a = *reinterpret_cast<volatile uint32_t*>(p.qthr + qa);
        const uint32_t ext_b = *reinterpret_cast<volatile uint32_t*>(p.qthr + qb);
        if (tid == 0) {
            tka.reset(ext_a);
            tkb.reset(ext_b);
        }
        __syncthreads();
        // a3: both tables for both queries
        if constexpr (DSUB != 0) {
            uint64_t rab[DSUB];
#pragma unroll
            for (int j = 0; j < DSUB; j++) rab[j] = res_ab[j * M + lm];
            const uint64_t negzero2 = p.negzero2;
#pragma unroll 2
            for (int i = 0; i < 16; i++) {
                const int c = lc0 + 16 * i;
                const float* pc = pq_t + static_cast<int64_t>(c) * (DSUB * M) + lm;
                float pv[DSUB];
#pragma unroll
                for (int j = 0; j < DSUB; j++) pv[j] = __ldg(pc + j * M);
                const uint64_t e = lut_entry_duo<DSUB>(pv, rab, negzero2);
                lut_e[c * 32 + lm] = e;
                lut_o[c * 32 + ((lm + 16) & 31)] = e;
            }
        } else {
            for (int i = 0; i < 16; i++) {
                const int c = lc0 + 16 * i;
                const float* pc = pq_t + static_cast<int64_t>(c) * dsub * M + lm;
                float a = 0.0f, b = 0.0f;
                for (int j = 0; j < dsub; j++) {
                    const float pj = __ldg(pc + j * M);
                    const uint64_t rr = res_ab[lm * dsub + j];
                    a = sqdiff_acc(a, __uint_as_float(static_cast<uint32_t>(rr)), pj);
                    b = sqdiff_acc(b, __uint_as_float(static_cast<uint32_t>(rr >> 32)), pj);
                }
                const uint64_t e = pack_f32x2(a, b);
                lut_e[c * 32 + lm] = e;
                lut_o[c * 32 + ((lm + 16) & 31)] = e;
            }
        }
        __syncthreads();
        buf ^= 1;
        if (tid == 0 && next_work < ngroups) duo_copy_group_async(&s_grp[buf], static_cast<const DuoGroup*>(p.groups) + next_work);

        // a4 + a5.  Iteration it: odd block (code it-1 . hi | code it . lo) finishes code it-1, even block walks code it.
        // it = 0 is the prologue (code -1 = zeros), it = nblk finishes the last codes and needs no even block.
        uint32_t thr_a = ext_a, thr_b = ext_b;
        uint64_t acc = 0ull;
        const uint32_t nblk = (n + kDuo32Threads - 1) / kDuo32Threads;
#define DUO32_ITER(PRV, CUR, LOADTO, U)                                                                   \
    {                                                                                                     \
        const uint32_t it = t0 + U;                                                                       \
        LOADTO = duo32_load_code(lp, (it + 1) * kDuo32Threads + tid, n);                                  \
        const uint64_t fin = duo32_odd(luto_b, PRV.hi, CUR.lo, ws2, ws1, bs, loff, keep2, cap2, acc);     \
        const uint32_t idx = it * kDuo32Threads + tid - kDuo32Threads; /* wraps for the prologue */       \
        const uint32_t ba = static_cast<uint32_t>(fin), bb = static_cast<uint32_t>(fin >> 32);            \
        const bool pa = idx < n && ba <= thr_a, pb = idx < n_b && bb <= thr_b;                            \
        if (__any_sync(0xffffffffu, pa || pb)) {                                                          \
            tka.push(pa, make_key(ba, idx));                                                              \
            tkb.push(pb, make_key(bb, idx));                                                              \
        }                                                                                                 \
        if (it < nblk) duo32_even(lute_b, CUR.lo, CUR.hi, ws2, ws1, bs, loff, keep0, acc);                \
    }
        uint32_t t0 = 0;
        for (; t0 + (kDuo32Unroll - 1) <= nblk; t0 += kDuo32Unroll) {
            DUO32_ITER(c0, c1, c2, 0)
            DUO32_ITER(c1, c2, c0, 1)
            DUO32_ITER(c2, c0, c1, 2)
            const int lim_a = thr_a == kInfBits ? 0 : kDuo32Cap - kDuo32Tile;
            const int lim_b = thr_b == kInfBits ? 0 : kDuo32Cap - kDuo32Tile;
            const int seen_a = *reinterpret_cast<volatile int*>(&tka.meta[1]);
            const int seen_b = *reinterpret_cast<volatile int*>(&tkb.meta[1]);
            if (__syncthreads_or(seen_a > lim_a || seen_b > lim_b)) {
                tka.flush<kDuo32Threads>(ext_a);
                tkb.flush<kDuo32Threads>(ext_b);
            }
            thr_a = tka.threshold();
            thr_b = tkb.threshold();
        }
        if (t0 <= nblk) {   // remaining 1..2 iterations (CTA-uniform); the queues have room for a whole tile
            DUO32_ITER(c0, c1, c2, 0)
            if (t0 + 1 <= nblk) DUO32_ITER(c1, c2, c0, 1)
        }
#undef DUO32_ITER
        __syncthreads();
        tka.flush<kDuo32Threads>(ext_a);
        tkb.flush<kDuo32Threads>(ext_b);
        {
            const int nb = tka.count();
            const uint64_t* s = tka.sorted();
            for (int i = tid; i < nb; i += kDuo32Threads) p.out_keys[static_cast<int64_t>(pair_a) * p.k + i] = s[i];
            if (tid == 0) {
                p.out_cnt[pair_a] = nb;
                if (nb == p.k) atomicMin(p.qthr + qa, static_cast<uint32_t>(s[p.k - 1] >> 32));
            }
        }
        if (has_b) {
            const int nb = tkb.count();
            const uint64_t* s = tkb.sorted();
            for (int i = tid; i < nb; i += kDuo32Threads) p.out_keys[static_cast<int64_t>(pair_b) * p.k + i] = s[i];
            if (tid == 0) {
                p.out_cnt[pair_b] = nb;
                if (nb == p.k) atomicMin(p.qthr + qb, static_cast<uint32_t>(s[p.k - 1] >> 32));
            }
        }
        // no barrier here: the one at the top of the next item separates these reads from its first writes
    }
}

template <int DSUB>
int launch_scan_duo32_t(const ScanParams& sp, const float* pq_t, int64_t npairs, int num_sms, cudaStream_t st) {
    size_t smem = duo32_smem_bytes(sp.d, sp.k);
    auto kernel = scan_duo32_kernel<DSUB>;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kDuo32Threads, smem) != cudaSuccess) return -1;
    if (per_sm < 1) return -1;
    int64_t grid = static_cast<int64_t>(per_sm) * num_sms;
    if (grid > npairs) grid = npairs;
    kernel<<<(unsigned)grid, kDuo32Threads, smem, st>>>(sp, pq_t);
    return cudaPeekAtLastError() == cudaSuccess ? 0 : -1;
}

// returns 0, -1 on a launch error (caller reads cudaGetLastError), -2 when the kernel does not fit (large k or d)
inline int launch_scan_duo32(const ScanParams& sp, const float* pq_t, int64_t npairs, int num_sms, cudaStream_t st) {
    if (duo32_smem_bytes(sp.d, sp.k) > 227 * 1024) return -2;
    switch (sp.dsub) {
        case 4: return launch_scan_duo32_t<4>(sp, pq_t, npairs, num_sms, st);
        case 8: return launch_scan_duo32_t<8>(sp, pq_t, npairs, num_sms, st);
        case 16: return launch_scan_duo32_t<16>(sp, pq_t, npairs, num_sms, st);
        default: return launch_scan_duo32_t<0>(sp, pq_t, npairs, num_sms, st);
    }
}

}  // namespace b200
