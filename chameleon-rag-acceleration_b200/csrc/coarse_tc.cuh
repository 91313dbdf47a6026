// coarse_tc.cuh -- K1 on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), SURVEY.md section 8a row a1.
//
// The coarse quantizer is the only dense contraction on the path.  The exact contract (oracle) is
// sum_j (q_j - c_j)^2 accumulated sequentially in fp32 without FMA, which no GEMM reproduces bit for bit, so the
// tensor cores are used as a PRE-FILTER that provably contains the answer:
//   1. approximate distance  s(q, c) = max(0, ||q||^2 + ||c||^2 - 2 q.c)  with q.c from ONE bf16 GEMM over a split operand:
//        q = qh + ql (+ 2^-18 |q|),  c = ch + cl (+ 2^-18 |c|),   q.c ~ qh.ch + qh.cl + ql.ch
//        A' = [qh | qh | ql]  (nq x 3d),  B' = [ch | cl | ch]  (nlist x 3d),  fp32 accumulation in TMEM;
//   2. the L smallest scores per query are candidates (coarse_select_kernel, L >= 2 nprobe);
//   3. coarse_rescore_kernel recomputes the candidates' distances in the oracle's exact form, takes the nprobe
//      smallest (distance, id), and proves that no non-candidate can belong to that set:
//        s_L - E  >  d_nprobe      (E = rigorous bound on |approximate - oracle| distance)
//      queries that fail the proof are flagged and re-done by the exact kernels (coarse_exact_flagged_kernel).
// The probed lists and their distances are therefore identical to the oracle's, bit for bit.
//
// GEMM kernel: persistent, warp-specialised (canonical sm_100 anatomy): warp 0 = TMA producer, warp 1 = MMA
// issuer (one thread, tcgen05.mma cta_group::1 kind::f16, M128 x N256 x K16), warp 2 = TMEM allocator, warps 4-7 =
// epilogue (tcgen05.ld 32x32b -> score -> global).  4-stage smem ring (A 128x64 + B 256x64 bf16, SWIZZLE_128B),
// 2 accumulator stages in TMEM (2 x 256 columns) so that the epilogue of tile i overlaps the MMAs of tile i+1.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace b200 {

constexpr int kTcBM = 128, kTcBN = 256, kTcBK = 64;
constexpr int kTcStages = 4;
constexpr int kTcThreads = 256;
constexpr uint32_t kTcABytes = kTcBM * kTcBK * 2;
constexpr uint32_t kTcBBytes = kTcBN * kTcBK * 2;
constexpr uint32_t kTcStageBytes = kTcABytes + kTcBBytes;
constexpr size_t kTcSmemBytes = 1024 + static_cast<size_t>(kTcStages) * kTcStageBytes + 256;

__host__ __device__ inline int tc_kpad(int d) { return ((3 * d + kTcBK - 1) / kTcBK) * kTcBK; }

// ---- operand preparation ---------------------------------------------------------------------------------------
// rows of x (n, d) f32 -> (n, kpad) bf16 as [hi | lo | hi] (centroids, lo_first = 0... see below) or [hi | hi | lo]
// (queries); also the squared norm in the oracle's sequential fp32 form.
//   queries:   [qh | qh | ql]      centroids: [ch | cl | ch]      =>   row-dot = qh.ch + qh.cl + ql.ch
__global__ void tc_split_rows_kernel(const float* __restrict__ x, int64_t n, int d, int kpad, int is_query,
                                     __nv_bfloat16* __restrict__ out, float* __restrict__ norms) {
    const int64_t row = blockIdx.x;
    const float* xr = x + row * d;
    __nv_bfloat16* o = out + row * kpad;
    for (int j = threadIdx.x; j < kpad; j += blockDim.x) {
        __nv_bfloat16 v = __float2bfloat16(0.0f);
        if (j < 3 * d) {
            const int seg = j / d, jj = j - seg * d;
            const float f = xr[jj];
            const __nv_bfloat16 hi = __float2bfloat16_rn(f);
            const __nv_bfloat16 lo = __float2bfloat16_rn(f - __bfloat162float(hi));
            const bool want_lo = is_query ? (seg == 2) : (seg == 1);
            v = want_lo ? lo : hi;
        }
        o[j] = v;
    }
    if (threadIdx.x == 0 && norms) {
        float acc = 0.0f;
        for (int j = 0; j < d; j++) acc = __fadd_rn(acc, __fmul_rn(xr[j], xr[j]));
        norms[row] = acc;
    }
}

__global__ void tc_max_kernel(const float* __restrict__ v, int64_t n, float* __restrict__ out) {
    __shared__ float s[256];
    float m = 0.0f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, v[i]);
    s[threadIdx.x] = m;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] = fmaxf(s[threadIdx.x], s[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = s[0];
}

// ---- PTX helpers -----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major operand tile in SWIZZLE_128B layout: rows of 128 B, 8-row atoms of 1024 B (SBO), version 1 (Blackwell)
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t addr) {
    uint64_t desc = 0;
    desc |= static_cast<uint64_t>((addr >> 4) & 0x3FFF);
    desc |= static_cast<uint64_t>(0) << 16;                  // leading byte offset: unused for swizzled K-major
    desc |= static_cast<uint64_t>(1024 >> 4) << 32;          // stride byte offset between 8-row atoms
    desc |= static_cast<uint64_t>(1) << 46;                  // descriptor version
    desc |= static_cast<uint64_t>(2) << 61;                  // SWIZZLE_128B
    return desc;
}
// kind::f16 instruction descriptor: D = F32, A = B = BF16, both K-major, M = 128, N = 256
__device__ __forceinline__ constexpr uint32_t tc_idesc() {
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(kTcBN >> 3) << 17) |
           (static_cast<uint32_t>(kTcBM >> 4) << 24);
}

struct TcGemmParams {
    const float* cnorm;   // (nlist)
    const float* qnorm;   // (nq)
    float* out;           // kTcScores: (nq, nlist) scores max(0, qnorm[q] + cnorm[c] - 2 q.c)  (approximate L2^2, >= 0
                          // so that the unsigned-key top-k orders them; clamping only moves towards the truth)
                          // kTcMinima: (nq, nchunks) minimum score of every 32-centroid chunk
    int64_t nq, nlist;
    int kblocks;          // kpad / 64
    int mtiles, ntiles;
    // kTcFilter: every centroid with score <= tau[q * tau_stride] is appended to cand[q * cap ..] (cand_cnt[q] counts
    // ALL of them, also those beyond cap)
    int64_t nchunks;
    const float* tau;
    int tau_stride;
    int32_t* cand;
    int* cand_cnt;
    int cap;
};

// The epilogue never needs the (nq, nlist) score matrix in HBM when the candidates are selected in two passes of the
// same deterministic GEMM:
//   pass 1 (kTcMinima)  minimum score of every 32-centroid chunk, (nq, nlist / 32);  the L-th smallest chunk minimum
//                       tau is an upper bound of the L-th smallest score (L chunks hold a score <= tau);
//   pass 2 (kTcFilter)  the same GEMM again; centroids with score <= tau are the candidates (at least L of them, a few
//                       more in expectation), every non-candidate has score > tau -- which is what the sufficiency
//                       proof of coarse_rescore_kernel needs.
// Recomputing 2 nq nlist 3d flops on the tensor cores is far cheaper than writing and re-reading the matrix
// (C3: 2.6 GB per 10k queries).
enum TcMode { kTcScores = 0, kTcMinima = 1, kTcFilter = 2 };

template <int MODE>
__global__ void __launch_bounds__(kTcThreads, 1)
coarse_tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const TcGemmParams p) {
    extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
    // 1024-byte aligned operand ring
    const uint32_t raw = smem_u32(tc_smem_raw);
    const uint32_t ring = (raw + 1023u) & ~1023u;
    unsigned char* ring_ptr = tc_smem_raw + (ring - raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring_ptr + static_cast<size_t>(kTcStages) * kTcStageBytes);
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kTcStages + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kTcStages + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kTcStages + 2 + s); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kTcStages + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles_total = p.mtiles * p.ntiles;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kTcStages; s++) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; s++) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < ntiles_total; t += gridDim.x) {
                const int m_blk = t % p.mtiles, n_blk = t / p.mtiles;
                for (int kb = 0; kb < p.kblocks; kb++) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t a_dst = ring + stage * kTcStageBytes;
                    const uint32_t b_dst = a_dst + kTcABytes;
                    mbar_expect_tx(full_bar(stage), kTcStageBytes);
                    tma_load_2d(a_dst, &tmA, kb * kTcBK, m_blk * kTcBM, full_bar(stage));
                    tma_load_2d(b_dst, &tmB, kb * kTcBK, n_blk * kTcBN, full_bar(stage));
                    if (++stage == kTcStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            constexpr uint32_t idesc = tc_idesc();
            for (int t = blockIdx.x; t < ntiles_total; t += gridDim.x) {
                mbar_wait(tempty_bar(as), aphase ^ 1u);     // epilogue has drained this accumulator stage
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * kTcBN);
                for (int kb = 0; kb < p.kblocks; kb++) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t a_addr = ring + stage * kTcStageBytes;
                    const uint64_t da = tc_smem_desc(a_addr), db = tc_smem_desc(a_addr + kTcABytes);
#pragma unroll
                    for (int k = 0; k < kTcBK / 16; k++) {
                        // advance 16 bf16 = 32 B inside the 128-byte swizzle span: +2 in the (addr >> 4) field
                        tc_mma_bf16(tmem_d, da + 2u * k, db + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    tc_commit(empty_bar(stage));             // frees the smem slot when these MMAs retire
                    if (++stage == kTcStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                tc_commit(tfull_bar(as));                    // accumulator complete -> epilogue
                if (++as == 2) {
                    as = 0;
                    aphase ^= 1u;
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> score -> global =====
        const int ew = warp & 3;                              // TMEM lane quarter this warp may access
        int as = 0;
        uint32_t aphase = 0;
        for (int t = blockIdx.x; t < ntiles_total; t += gridDim.x) {
            const int m_blk = t % p.mtiles, n_blk = t / p.mtiles;
            mbar_wait(tfull_bar(as), aphase);
            tc_fence_after();
            const int64_t row = static_cast<int64_t>(m_blk) * kTcBM + ew * 32 + lane;
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(as * kTcBN);
            float* orow = p.out + row * (MODE == kTcMinima ? p.nchunks : p.nlist);
            const float qn = row < p.nq ? __ldg(p.qnorm + row) : 0.0f;
            float tau = 0.0f;
            if (MODE == kTcFilter && row < p.nq) tau = __ldg(p.tau + row * p.tau_stride);
            uint32_t hit[kTcBN / 32];   // kTcFilter: bit j of hit[i] = column 32 i + j of this tile is a candidate
#pragma unroll
            for (int i = 0; i < kTcBN / 32; i++) hit[i] = 0u;
#pragma unroll 1
            for (int c0 = 0; c0 < kTcBN; c0 += 32) {
                uint32_t v[32];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
                      "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
                      "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
                      "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr + static_cast<uint32_t>(c0)));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const int64_t cbase = static_cast<int64_t>(n_blk) * kTcBN + c0;
                if (row < p.nq && cbase < p.nlist) {
                    const bool full = cbase + 32 <= p.nlist && (p.nlist & 3) == 0;
                    if (MODE == kTcScores) {
                        if (full) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 cn = __ldg(reinterpret_cast<const float4*>(p.cnorm + cbase + j));
                                float4 o;
                                o.x = fmaxf(0.0f, fmaf(-2.0f, __uint_as_float(v[j + 0]), cn.x + qn));
                                o.y = fmaxf(0.0f, fmaf(-2.0f, __uint_as_float(v[j + 1]), cn.y + qn));
                                o.z = fmaxf(0.0f, fmaf(-2.0f, __uint_as_float(v[j + 2]), cn.z + qn));
                                o.w = fmaxf(0.0f, fmaf(-2.0f, __uint_as_float(v[j + 3]), cn.w + qn));
                                *reinterpret_cast<float4*>(orow + cbase + j) = o;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; j++)
                                if (cbase + j < p.nlist)
                                    orow[cbase + j] = fmaxf(
                                        0.0f, fmaf(-2.0f, __uint_as_float(v[j]), __ldg(p.cnorm + cbase + j) + qn));
                        }
                    } else {
                        // unclamped scores of the chunk (same expression as above, so the three modes agree bit for
                        // bit); columns past nlist count as +inf
                        float sc[32];
                        if (full) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 cn = __ldg(reinterpret_cast<const float4*>(p.cnorm + cbase + j));
                                sc[j + 0] = fmaf(-2.0f, __uint_as_float(v[j + 0]), cn.x + qn);
                                sc[j + 1] = fmaf(-2.0f, __uint_as_float(v[j + 1]), cn.y + qn);
                                sc[j + 2] = fmaf(-2.0f, __uint_as_float(v[j + 2]), cn.z + qn);
                                sc[j + 3] = fmaf(-2.0f, __uint_as_float(v[j + 3]), cn.w + qn);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; j++)
                                sc[j] = cbase + j < p.nlist
                                            ? fmaf(-2.0f, __uint_as_float(v[j]), __ldg(p.cnorm + cbase + j) + qn)
                                            : __uint_as_float(kInfBits);
                        }
                        float mn = sc[0];
#pragma unroll
                        for (int j = 1; j < 32; j++) mn = fminf(mn, sc[j]);
                        mn = fmaxf(0.0f, mn);
                        if (MODE == kTcMinima) {
                            orow[cbase >> 5] = mn;
                        } else {
                            // clamping commutes with the comparison (tau >= 0)
                            uint32_t m = 0u;
#pragma unroll
                            for (int j = 0; j < 32; j++) m |= (sc[j] <= tau ? 1u : 0u) << j;
#pragma unroll
                            for (int i = 0; i < kTcBN / 32; i++)
                                if (i == c0 / 32) hit[i] = m;
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(tempty_bar(as));
            if (++as == 2) {
                as = 0;
                aphase ^= 1u;
            }
            if (MODE == kTcFilter) {
                // the accumulator stage is already released; one returning atomic per row and tile (its latency would
                // otherwise be paid per candidate)
                int total = 0;
#pragma unroll
                for (int i = 0; i < kTcBN / 32; i++) total += __popc(hit[i]);
                if (total > 0) {
                    int pos = atomicAdd(p.cand_cnt + row, total);
#pragma unroll
                    for (int i = 0; i < kTcBN / 32; i++) {
                        uint32_t m = hit[i];
                        while (m) {
                            const int j = __ffs(m) - 1;
                            m &= m - 1;
                            if (pos < p.cap)
                                p.cand[row * p.cap + pos] =
                                    static_cast<int32_t>(static_cast<int64_t>(n_blk) * kTcBN + 32 * i + j);
                            pos++;
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- exact rescoring of the candidates + proof of sufficiency ---------------------------------------------------
// One CTA per query.  cand: (nq, cap) centroid ids; the first min(cand_cnt[q], cap) are valid (cand_cnt == nullptr:
// all cap).  bound[q * bound_stride] is a score every NON-candidate is known to reach or exceed: the L-th smallest
// approximate score (radix-select path) or the filter threshold tau (two-pass path).  Writes the nprobe smallest
// (exact distance, id) and flags the query when the candidates cannot be proven sufficient (or overflowed cap).
// eps_rel * ||q|| * cmax + eps_abs * (||q|| + cmax)^2 bounds |approximate - oracle| distance.
constexpr int kRescoreThreads = 128;
constexpr int kRescoreJ = 32;                       // dimensions staged per step
constexpr int kRescoreTileWords = kRescoreThreads * (kRescoreJ + 1);

__host__ __device__ inline size_t rescore_smem_bytes(int d, int nprobe) {
    return sizeof(float) * (static_cast<size_t>((d + 3) & ~3) + kRescoreTileWords) + TopK::smem_bytes(nprobe, 2048) +
           sizeof(int) * kRescoreThreads;
}

__global__ void __launch_bounds__(kRescoreThreads)
coarse_rescore_kernel(const float* __restrict__ xq, const float* __restrict__ cent, const float* __restrict__ qnorm,
                      const float* __restrict__ cmax2, const int32_t* __restrict__ cand,
                      const int* __restrict__ cand_cnt, int cap, const float* __restrict__ bound, int bound_stride,
                      int d, int64_t nlist, int nprobe, float eps_rel, float eps_abs, int32_t* __restrict__ probe32,
                      int64_t* __restrict__ ids64, float* __restrict__ dis_out, int* __restrict__ flags) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* sq = reinterpret_cast<float*>(smem_raw);                 // query, d floats
    float* tile = sq + ((d + 3) & ~3);                              // [128 candidates][32 + 1] staged centroid slices
    TopK tk;
    tk.bind(tile + kRescoreTileWords, nprobe, 2048);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t q = blockIdx.x;
    const int found = cand_cnt ? cand_cnt[q] : cap;
    const int L = found < cap ? found : cap;
    for (int j = tid; j < d; j += kRescoreThreads) sq[j] = xq[q * d + j];
    if (tid == 0) tk.reset(kInfBits);
    __syncthreads();
    uint32_t thr = kInfBits;
    // 128 candidates at a time, one per thread.  Each thread must add its candidate's d terms in order, so the rows
    // are staged through shared memory in 32-dimension slices: the warps read them from global memory coalesced
    // (lane = dimension) and every thread then walks its own row (stride 33 words: conflict-free).
    int* s_ids = reinterpret_cast<int*>(tk.meta + 4);               // [128] candidate ids of the current batch
    const bool vec4 = (d & 3) == 0;
    for (int base = 0; base < L; base += kRescoreThreads) {
        const int nc = min(kRescoreThreads, L - base);
        const int32_t my_id = tid < nc ? cand[q * cap + base + tid] : -1;
        s_ids[tid] = my_id;
        __syncthreads();
        float acc = 0.0f;
        for (int j0 = 0; j0 < d; j0 += kRescoreJ) {
            const int jn = min(kRescoreJ, d - j0);
            if (vec4) {
                // 8 threads x float4 cover one 128-byte row slice, 16 rows per pass; all loads of a slice in flight
                const int c4 = (tid & 7) * 4, r0 = tid >> 3;
                float4 v[kRescoreThreads / 16];
#pragma unroll
                for (int ps = 0; ps < kRescoreThreads / 16; ps++) {
                    const int i = r0 + 16 * ps;
                    const int32_t id = i < nc ? s_ids[i] : -1;
                    v[ps] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    if (id >= 0 && c4 < jn)
                        v[ps] = __ldg(reinterpret_cast<const float4*>(cent + static_cast<int64_t>(id) * d + j0 + c4));
                }
#pragma unroll
                for (int ps = 0; ps < kRescoreThreads / 16; ps++) {
                    float* t = tile + (r0 + 16 * ps) * (kRescoreJ + 1) + c4;
                    t[0] = v[ps].x;
                    t[1] = v[ps].y;
                    t[2] = v[ps].z;
                    t[3] = v[ps].w;
                }
            } else {
                for (int i = warp; i < nc; i += kRescoreThreads / 32) {
                    const int32_t id = s_ids[i];
                    if (lane < jn && id >= 0)
                        tile[i * (kRescoreJ + 1) + lane] = __ldg(cent + static_cast<int64_t>(id) * d + j0 + lane);
                }
            }
            __syncthreads();
            if (my_id >= 0) {
                const float* row = tile + tid * (kRescoreJ + 1);
                for (int jj = 0; jj < jn; jj++) acc = sqdiff_acc(acc, sq[j0 + jj], row[jj]);
            }
            __syncthreads();
        }
        const uint32_t bits = my_id >= 0 ? __float_as_uint(acc) : 0xffffffffu;
        tk.push(bits <= thr, make_key(bits, static_cast<uint32_t>(my_id)));
        tk.sync_and_flush_if_over<kRescoreThreads>(2048 - kRescoreThreads, kInfBits);
        thr = tk.threshold();
    }
    __syncthreads();
    tk.flush<kRescoreThreads>(kInfBits);
    const int nb = tk.count();
    const uint64_t* s = tk.sorted();
    for (int i = tid; i < nprobe; i += kRescoreThreads) {
        int32_t id = -1;
        float dv = FLT_MAX;
        if (i < nb) {
            id = static_cast<int32_t>(s[i] & 0xffffffffu);
            dv = __uint_as_float(static_cast<uint32_t>(s[i] >> 32));
        }
        if (probe32) probe32[q * nprobe + i] = id;
        if (ids64) ids64[q * nprobe + i] = id;
        if (dis_out) dis_out[q * nprobe + i] = dv;
    }
    if (tid == 0) {
        int flag = found > cap ? 1 : 0;
        if (found < nlist) {
            // every non-candidate has an approximate score >= sL
            const float sL = bound[q * bound_stride];
            const float qn = qnorm[q];
            const float a = sqrtf(qn), b = sqrtf(*cmax2);
            const float E = eps_rel * a * b + eps_abs * (a + b) * (a + b);
            const float dk = nb >= nprobe ? __uint_as_float(static_cast<uint32_t>(s[nprobe - 1] >> 32)) : FLT_MAX;
            // non-candidate oracle distance >= sL - E ; need it strictly above the nprobe-th exact distance
            if (!(sL - E > dk)) flag = 1;
        }
        flags[q] = flag;
    }
}

// exact K1 for the flagged queries only: one CTA per query, exits at once when the flag is clear.
__global__ void __launch_bounds__(kThreads)
coarse_exact_flagged_kernel(const float* __restrict__ xq, const float* __restrict__ cent, const int* __restrict__ flags,
                            int d, int64_t nlist, int nprobe, int32_t* __restrict__ probe32,
                            int64_t* __restrict__ ids64, float* __restrict__ dis_out, int* __restrict__ nflagged) {
    const int64_t q = blockIdx.x;
    if (!flags[q]) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* sq = reinterpret_cast<float*>(smem_raw);
    TopK tk;
    tk.bind(sq + ((d + 3) & ~3), nprobe, kSelCap);
    const int tid = threadIdx.x;
    for (int j = tid; j < d; j += kThreads) sq[j] = xq[q * d + j];
    if (tid == 0) {
        tk.reset(kInfBits);
        atomicAdd(nflagged, 1);
    }
    __syncthreads();
    uint32_t thr = kInfBits;
    for (int64_t base = 0; base < nlist; base += kSelTile) {
#pragma unroll 1
        for (int u = 0; u < kSelTile / kThreads; u++) {
            const int64_t c = base + u * kThreads + tid;
            uint32_t bits = 0xffffffffu;
            if (c < nlist) {
                const float* cp = cent + c * d;
                float acc = 0.0f;
                for (int j = 0; j < d; j++) acc = sqdiff_acc(acc, sq[j], __ldg(cp + j));
                bits = __float_as_uint(acc);
            }
            tk.push(bits <= thr, make_key(bits, static_cast<uint32_t>(c)));
        }
        tk.sync_and_flush_if_over<kThreads>(kSelCap - kSelTile, kInfBits);
        thr = tk.threshold();
    }
    __syncthreads();
    tk.flush<kThreads>(kInfBits);
    const int nb = tk.count();
    const uint64_t* s = tk.sorted();
    for (int i = tid; i < nprobe; i += kThreads) {
        int32_t id = -1;
        float dv = FLT_MAX;
        if (i < nb) {
            id = static_cast<int32_t>(s[i] & 0xffffffffu);
            dv = __uint_as_float(static_cast<uint32_t>(s[i] >> 32));
        }
        if (probe32) probe32[q * nprobe + i] = id;
        if (ids64) ids64[q * nprobe + i] = id;
        if (dis_out) dis_out[q * nprobe + i] = dv;
    }
}

// ---- host side: tensor maps --------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled tc_get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// (rows, kpad) bf16 row-major, box = (64 elements, box_rows rows), 128-byte swizzle, zero fill out of bounds
inline bool tc_make_map(CUtensorMap* map, const void* base, int64_t rows, int kpad, int box_rows) {
    PFN_encodeTiled enc = tc_get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(kpad), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(kpad) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(kTcBK), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

}  // namespace b200
