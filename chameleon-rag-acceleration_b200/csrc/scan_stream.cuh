// scan_stream.cuh -- the per-query-table filter (scan_qlut.cuh) as a STREAMING pipeline: no top-k inside the scan.
//
// Why.  scan_qlut_kernel evaluates its survivors and folds its top-k lists inside every work item: a chain of
// dependent phases (queue -> barrier -> a few lanes walk the codebook -> barrier -> fold -> barrier -> publish) that
// costs ~10 us per item whatever the list length.  On an 8-way vector shard (1.5k-code lists) that chain, not the
// scan, is the run time (ncu: 35 % issue slots, barrier and long-scoreboard stalls on top).
//
// How.  Everything that was serial inside an item becomes its own fully parallel kernel:
//   S  st_boot_kernel    one CTA per query: scan positions (prefix of list sizes in probe order), ||q - c||^2 of every
//                        probed list, and a BOOTSTRAP THRESHOLD: of the first <= 2048 codes in scan order (the nearest
//                        lists), the candidates with the smallest estimated distance are evaluated exactly; their k-th
//                        smallest exact distance bounds the query's final k-th distance from above;
//   A  st_filter_kernel  persistent CTAs over (list, <= 4 queries) work items: copy the queries' tables, scan the list
//                        with the FIXED thresholds, append survivors to a global record buffer.  Two barriers per item,
//                        no top-k state, no exact arithmetic; every warp owns a 64-record chunk at a time, so an append
//                        is a coalesced store and one global atomic per 64 survivors;
//   B  st_eval_kernel    one thread per survivor: exact distance in the oracle's operation order (residual formed on
//                        the fly), kept if <= the threshold: key = (distance bits << 32) | scan position, appended to
//                        the query's slab;
//   C  st_select_kernel  one CTA per query: the k smallest keys of the slab = the oracle's (distance, probe rank,
//                        offset) order; id lookup.
// The result is exactly the oracle's: the threshold is the k-th smallest of k real distances (so the true top-k lies at
// or below it), the filter never drops a code whose exact distance is <= the threshold (scan_qlut.cuh's bound), and every
// kept code is evaluated exactly.
// Buffers are sized for the typical survivor rate.  If the record buffer overflows (degenerate data, fewer than k
// vectors in the probed lists) a flag is raised on the device and the guarded launches that follow -- scan_qlut_kernel
// and merge_query_kernel over the same work items -- recompute the batch; if only a query's slab overflows (hundreds of
// duplicate codes at its k-th distance) they recompute that query alone; they return at once otherwise.
#pragma once
#include "scan_qlut.cuh"

namespace b200 {

constexpr int kStChunk = 64;          // survivor records per chunk
constexpr int kStBootCodes = 2048;    // codes (in scan order) the bootstrap looks at
constexpr int kStBootThreads = 256;
constexpr int kStSelCap = 2048;
constexpr int kStSelThreads = 128;

struct StCounters {
    unsigned int nchunks;             // chunks handed out
    int overflow;                     // != 0: the fallback launches recompute the batch (1: scan positions exceed 32 bits,
                                      // 2: record buffer full, 4: a query's slab full)
    unsigned long long records;       // statistics: survivor records, exact evaluations
    unsigned long long evals;
    unsigned long long kept;          // keys admitted to the slabs
    unsigned int maxkept;             // largest slab
    unsigned int nflag;               // queries answered by the fallback launches
};

struct StParams {
    uint2* srec;                      // (max_chunks * 64) survivor records: (work item, (offset << 4) | query mask)
    unsigned int* sfill;              // (max_chunks) records in each chunk
    unsigned int max_chunks;
    StCounters* ctr;
    uint64_t* slab;                   // (nq, capq) keys of the codes at or below the query's threshold
    unsigned int* qcnt;               // (nq)
    int* qflag;                       // (nq) != 0: the query's slab overflowed, the fallback launches answer it
    uint64_t* qkey;                   // (nq) bootstrap threshold as a full key (distance bits, scan position): keys are
                                      // distinct, so duplicate codes at the k-th distance cannot flood a slab
    int capq;
    uint32_t* prefix;                 // (nq, nprobe) scan position of the first code of every probed list
    float* pdis;                      // (nq, nprobe) ||fl(q - c)||^2
    const int64_t* ids;
    int boot_codes;                   // candidates the bootstrap looks at (<= kStBootCodes)
    int gsz;                          // pairs per work item: 4, or 2 (the two-query filters)
    int two_kind;                     // gsz == 2: 1 = st_filter_kernel<16, true>, 2 = st_filter2_kernel (bulk-async tiles)
    float* D;                         // (nq, k)
    int64_t* I;
};

__host__ __device__ inline int st_boot_nsel(int k) {
    const int n = 4 * k;
    return n < 32 ? 32 : n > 1024 ? 1024 : n;
}
__host__ __device__ inline size_t st_boot_smem(int d, int M, int nprobe, int k) {
    // query | the query's table | list sizes | prefix | list ids | TopK
    return sizeof(float) * ((d + 3) & ~3) + sizeof(uint16_t) * 256 * (M + 2) + 3 * sizeof(uint32_t) * nprobe +
           TopK::smem_bytes(st_boot_nsel(k), kStSelCap) + 64;
}

// exact distance with the residual formed on the fly: r_j = fl(q_j - c_j), then the oracle's sums
template <int M>
__device__ __forceinline__ float st_exact(const uint8_t* __restrict__ code, const float* __restrict__ q,
                                          const float* __restrict__ c, const float* __restrict__ pq, int dsub) {
    float acc = 0.0f;
#pragma unroll 1
    for (int m0 = 0; m0 < M; m0 += 4) {
        const uint32_t cw = __ldg(reinterpret_cast<const uint32_t*>(code + m0));
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int m = m0 + b;
            const uint32_t cv = (cw >> (8 * b)) & 255u;
            const float* pc = pq + (static_cast<int64_t>(m) * 256 + cv) * dsub;
            const float* qq = q + m * dsub;
            const float* cc = c + m * dsub;
            float t = 0.0f;
            if ((dsub & 3) == 0) {
                for (int j = 0; j < dsub; j += 4) {
                    const float4 pv = __ldg(reinterpret_cast<const float4*>(pc + j));
                    const float4 qv = *reinterpret_cast<const float4*>(qq + j);
                    const float4 cv4 = __ldg(reinterpret_cast<const float4*>(cc + j));
                    t = sqdiff_acc(t, __fsub_rn(qv.x, cv4.x), pv.x);
                    t = sqdiff_acc(t, __fsub_rn(qv.y, cv4.y), pv.y);
                    t = sqdiff_acc(t, __fsub_rn(qv.z, cv4.z), pv.z);
                    t = sqdiff_acc(t, __fsub_rn(qv.w, cv4.w), pv.w);
                }
            } else {
                for (int j = 0; j < dsub; j++) t = sqdiff_acc(t, __fsub_rn(qq[j], __ldg(cc + j)), __ldg(pc + j));
            }
            acc = __fadd_rn(acc, t);
        }
    }
    return acc;
}

// ------------------------------------------------------------------------------------------------------------------
// S: per query -- scan positions, coarse distances, bootstrap threshold
// ------------------------------------------------------------------------------------------------------------------
template <int M>
__global__ void __launch_bounds__(kStBootThreads)
st_boot_kernel(const float* __restrict__ xq, const float* __restrict__ cent, const float* __restrict__ pq,
               const int64_t* __restrict__ offsets, const uint8_t* __restrict__ codes,
               const int32_t* __restrict__ probe, int nprobe, int d, int dsub, int k, QlParams ql, StParams st,
               uint32_t* __restrict__ qthr, int64_t boot_lo, int64_t boot_hi) {
    extern __shared__ __align__(16) unsigned char smem_boot[];
    const int dpad = (d + 3) & ~3;
    float* qv = reinterpret_cast<float*>(smem_boot);
    // the query's table, rows padded to M + 2 entries: an odd number of 32-bit words, so that the gathers of a warp
    // (same m, 32 different code values) spread over the banks
    constexpr int kRow = M + 2;
    uint16_t* s_lut = reinterpret_cast<uint16_t*>(qv + dpad);                 // [256][M + 2]
    uint32_t* s_sz = reinterpret_cast<uint32_t*>(s_lut + 256 * kRow);
    uint32_t* s_pre = s_sz + nprobe;
    int32_t* s_list = reinterpret_cast<int32_t*>(s_pre + nprobe);
    const int nsel = st_boot_nsel(k);
    uintptr_t tkp = (reinterpret_cast<uintptr_t>(s_list + nprobe) + 7) & ~static_cast<uintptr_t>(7);
    TopK tk;
    tk.bind(reinterpret_cast<void*>(tkp), nsel, kStSelCap);
    __shared__ unsigned long long s_total;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t q = blockIdx.x;
    for (int j = tid; j < d; j += kStBootThreads) qv[j] = xq[q * d + j];
    if (q >= boot_lo && q < boot_hi) {   // (the other queries only need their scan positions and coarse distances)
        const uint32_t* src = reinterpret_cast<const uint32_t*>(ql.qlut + q * 256 * M);
        uint32_t* dst = reinterpret_cast<uint32_t*>(s_lut);
        for (int i = tid; i < 256 * M / 2; i += kStBootThreads) dst[(i / (M / 2)) * (kRow / 2) + (i % (M / 2))] = __ldg(src + i);
    }
    for (int r = tid; r < nprobe; r += kStBootThreads) {
        const int l = probe[q * nprobe + r];
        int64_t sz = 0;
        if (l >= 0) sz = offsets[l + 1] - offsets[l];
        s_list[r] = sz > 0 ? l : -1;
        s_sz[r] = static_cast<uint32_t>(sz > 0 ? sz : 0);
    }
    if (tid == 0) tk.reset(kInfBits);
    __syncthreads();
    // scan positions: exclusive prefix of the list sizes in probe order (warp 0, 32 ranks at a time)
    if (wid == 0) {
        unsigned long long carry = 0ull;
        for (int r0 = 0; r0 < nprobe; r0 += 32) {
            const int r = r0 + lane;
            const unsigned long long v = r < nprobe ? s_sz[r] : 0u;
            unsigned long long x = v;
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
            }
            if (r < nprobe) {
                const unsigned long long pre = carry + x - v;
                s_pre[r] = static_cast<uint32_t>(pre);
                st.prefix[q * nprobe + r] = static_cast<uint32_t>(pre);
            }
            carry += __shfl_sync(0xffffffffu, x, 31);
        }
        if (lane == 0) {
            s_total = carry;
            if (carry >= (1ull << 32)) atomicOr(&st.ctr->overflow, 1);   // scan positions do not fit 32 bits: fallback
            st.qcnt[q] = 0u;
            st.qflag[q] = 0;
        }
    }
    // coarse distances of every probed list (any summation order: only a bound is needed)
    for (int r = wid; r < nprobe; r += kStBootThreads / 32) {
        const int l = s_list[r];
        float a = 0.0f;
        if (l >= 0) {
            for (int j = lane; j < d; j += 32) {
                const float rj = __fsub_rn(qv[j], __ldg(cent + static_cast<int64_t>(l) * d + j));
                a = fmaf(rj, rj, a);
            }
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        }
        if (lane == 0) st.pdis[q * nprobe + r] = a;
    }
    __syncthreads();
    if (q < boot_lo || q >= boot_hi) {
        // multi-GPU: another rank bootstraps this query's threshold on ITS shard (any shard's k-th best distance bounds
        // the global one); the thresholds are exchanged before the filter runs (st_apply_thresholds_kernel)
        if (tid == 0) {
            qthr[q] = kInfBits;
            st.qkey[q] = kPadKey;
        }
        return;
    }
    // candidates: the first codes in scan order, ranked by the estimated distance dis0 + SB + sum_m A; all of them
    // go to the queue (it holds kStBootCodes keys), one fold at the end
    const uint32_t ncand = static_cast<uint32_t>(s_total < static_cast<unsigned long long>(st.boot_codes) ? s_total : st.boot_codes);
    const float sq = ql.qscale[q], aq = ql.qamin[q];
    const float inv = sq > 0.0f ? 1.0f / sq : 0.0f;
    constexpr int kPer = kStBootCodes / kStBootThreads;   // 8 candidates per thread
    constexpr int kBatch = 2;   // small batches: after the first fold most candidates fail the comparison with bthr
    uint32_t bthr = kInfBits;
#pragma unroll 1
    for (int i0 = 0; i0 < kPer; i0 += kBatch) {
        uint4 cv[kBatch][M / 16];
        float add[kBatch];
#pragma unroll
        for (int i = 0; i < kBatch; i++) {
            const uint32_t pos = (i0 + i) * kStBootThreads + tid;
            add[i] = 0.0f;
#pragma unroll
            for (int h = 0; h < M / 16; h++) cv[i][h] = make_uint4(0u, 0u, 0u, 0u);
            if (pos < ncand) {
                int r = 0;
                while (r + 1 < nprobe && s_pre[r + 1] <= pos) r++;       // last rank whose first position is <= pos
                const int l = s_list[r];
                const int64_t row = offsets[l] + (pos - s_pre[r]);
                const uint4* cp = reinterpret_cast<const uint4*>(codes + row * M);
#pragma unroll
                for (int h = 0; h < M / 16; h++) cv[i][h] = __ldg(cp + h);
                add[i] = st.pdis[q * nprobe + r] + __ldg(ql.sbmin + l) +
                         __ldg(ql.sbstep + l) * static_cast<float>(__ldg(ql.snorm + row)) + aq;
            }
        }
#pragma unroll
        for (int i = 0; i < kBatch; i++) {
            const uint32_t pos = (i0 + i) * kStBootThreads + tid;
            uint32_t u = 0u;
#pragma unroll
            for (int h = 0; h < M / 16; h++) {
                const uint32_t w[4] = {cv[i][h].x, cv[i][h].y, cv[i][h].z, cv[i][h].w};
#pragma unroll
                for (int mm = 0; mm < 16; mm++)
                    u += s_lut[((w[mm >> 2] >> (8 * (mm & 3))) & 255u) * kRow + 16 * h + mm];
            }
            const float est = add[i] + static_cast<float>(u) * inv;
            uint32_t bits = __float_as_uint(fmaxf(est, 0.0f));
            if (!(est == est)) bits = 0x7f7fffffu;                    // NaN estimates rank last
            tk.push(pos < ncand && bits <= bthr, make_key(bits, pos));
        }
        // fold between the batches: the later candidates are compared with the nsel-th estimate found so far
        tk.sync_and_flush_if_over<kStBootThreads>(kStSelCap - kBatch * kStBootThreads, kInfBits);
        bthr = tk.threshold();
    }
    __syncthreads();
    tk.flush<kStBootThreads>(kInfBits);
    // exact distances of the selected candidates; their k-th smallest is the threshold
    const int ns = tk.count();
    uint64_t mine[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const int i = tid + u * kStBootThreads;
        mine[u] = kPadKey;
        if (i < ns) {
            const uint32_t pos = static_cast<uint32_t>(tk.sorted()[i] & 0xffffffffu);
            int r = 0;
            while (r + 1 < nprobe && s_pre[r + 1] <= pos) r++;
            const int l = s_list[r];
            const int64_t row = offsets[l] + (pos - s_pre[r]);
            const uint32_t b = __float_as_uint(st_exact<M>(codes + row * M, qv, cent + static_cast<int64_t>(l) * d, pq, dsub));
            mine[u] = make_key(b, pos);
        }
    }
    __syncthreads();
    if (tid == 0) tk.reset(kInfBits);
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 4; u++) tk.push(mine[u] != kPadKey, mine[u]);
    __syncthreads();
    tk.flush<kStBootThreads>(kInfBits);
    if (tid == 0) {
        uint32_t t = kInfBits;
        uint64_t tkey = kPadKey;
        if (tk.count() >= k) {
            const uint64_t key = tk.sorted()[k - 1];
            if (static_cast<uint32_t>(key >> 32) < kInfBits) {
                t = static_cast<uint32_t>(key >> 32);
                tkey = key;
            }
        }
        qthr[q] = t;
        st.qkey[q] = tkey;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// S, fast form for k <= 4 W (W = 8, 16, 32 winners per warp): no block-level selection at all.  Every warp ranks its
// 256 candidates by the estimate, keeps its W best (W rounds of a warp arg-min), evaluates them exactly with 16 lanes
// per candidate (lane = sub-quantizer; the 16 partial sums are then added one after the other in the oracle's order,
// so the distance is the oracle's bit for bit), and warp 0 sorts the 8 W exact keys: their k-th smallest is the
// threshold.  Three barriers per query.
// ------------------------------------------------------------------------------------------------------------------
template <int M>
__host__ __device__ inline size_t st_bootw_smem(int d, int nprobe, int W) {
    return sizeof(float) * ((d + 3) & ~3) + sizeof(uint16_t) * 256 * (M + 2) + 3 * sizeof(uint32_t) * nprobe +
           sizeof(int64_t) * nprobe + sizeof(uint64_t) * 8 * W + 64;
}

template <int M, int W>
__global__ void __launch_bounds__(kStBootThreads)
st_boot_warp_kernel(const float* __restrict__ xq, const float* __restrict__ cent, const float* __restrict__ pq,
                    const int64_t* __restrict__ offsets, const uint8_t* __restrict__ codes,
                    const int32_t* __restrict__ probe, int nprobe, int d, int dsub, int k, QlParams ql, StParams st,
                    uint32_t* __restrict__ qthr, int64_t boot_lo, int64_t boot_hi) {
    static_assert(M % 16 == 0, "16 lanes per candidate, M / 16 sub-quantizers per lane");
    extern __shared__ __align__(16) unsigned char smem_boot[];
    const int dpad = (d + 3) & ~3;
    constexpr int kRow = M + 2;
    float* qv = reinterpret_cast<float*>(smem_boot);
    uint16_t* s_lut = reinterpret_cast<uint16_t*>(qv + dpad);                 // [256][M + 2]
    uint32_t* s_sz = reinterpret_cast<uint32_t*>(s_lut + 256 * kRow);
    uint32_t* s_pre = s_sz + nprobe;
    int32_t* s_list = reinterpret_cast<int32_t*>(s_pre + nprobe);
    uintptr_t al = (reinterpret_cast<uintptr_t>(s_list + nprobe) + 7) & ~static_cast<uintptr_t>(7);
    int64_t* s_off = reinterpret_cast<int64_t*>(al);                          // first row of every probed list
    uint64_t* s_keys = reinterpret_cast<uint64_t*>(s_off + nprobe);           // [8 W] exact keys
    __shared__ unsigned long long s_total;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t q = blockIdx.x;
    for (int j = tid; j < d; j += kStBootThreads) qv[j] = xq[q * d + j];
    if (q >= boot_lo && q < boot_hi) {   // (the other queries only need their scan positions and coarse distances)
        const uint32_t* src = reinterpret_cast<const uint32_t*>(ql.qlut + q * 256 * M);
        uint32_t* dst = reinterpret_cast<uint32_t*>(s_lut);
        for (int i = tid; i < 256 * M / 2; i += kStBootThreads) dst[(i / (M / 2)) * (kRow / 2) + (i % (M / 2))] = __ldg(src + i);
    }
    for (int r = tid; r < nprobe; r += kStBootThreads) {
        const int l = probe[q * nprobe + r];
        int64_t sz = 0, beg = 0;
        if (l >= 0) {
            beg = offsets[l];
            sz = offsets[l + 1] - beg;
        }
        s_list[r] = sz > 0 ? l : -1;
        s_sz[r] = static_cast<uint32_t>(sz > 0 ? sz : 0);
        s_off[r] = beg;
    }
    __syncthreads();
    if (wid == 0) {
        unsigned long long carry = 0ull;
        for (int r0 = 0; r0 < nprobe; r0 += 32) {
            const int r = r0 + lane;
            const unsigned long long v = r < nprobe ? s_sz[r] : 0u;
            unsigned long long x = v;
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
            }
            if (r < nprobe) {
                const unsigned long long pre = carry + x - v;
                s_pre[r] = static_cast<uint32_t>(pre);
                st.prefix[q * nprobe + r] = static_cast<uint32_t>(pre);
            }
            carry += __shfl_sync(0xffffffffu, x, 31);
        }
        if (lane == 0) {
            s_total = carry;
            if (carry >= (1ull << 32)) atomicOr(&st.ctr->overflow, 1);
            st.qcnt[q] = 0u;
            st.qflag[q] = 0;
        }
    }
    for (int r = wid; r < nprobe; r += kStBootThreads / 32) {
        const int l = s_list[r];
        float a = 0.0f;
        if (l >= 0) {
            for (int j = lane; j < d; j += 32) {
                const float rj = __fsub_rn(qv[j], __ldg(cent + static_cast<int64_t>(l) * d + j));
                a = fmaf(rj, rj, a);
            }
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        }
        if (lane == 0) st.pdis[q * nprobe + r] = a;
    }
    __syncthreads();
    if (q < boot_lo || q >= boot_hi) {   // see st_boot_kernel
        if (tid == 0) {
            qthr[q] = kInfBits;
            st.qkey[q] = kPadKey;
        }
        return;
    }
    const uint32_t ncand = static_cast<uint32_t>(s_total < static_cast<unsigned long long>(st.boot_codes) ? s_total : st.boot_codes);
    const float sq = ql.qscale[q], aq = ql.qamin[q];
    const float inv = sq > 0.0f ? 1.0f / sq : 0.0f;
    constexpr int kPer = kStBootCodes / kStBootThreads;   // 8 candidates per lane: positions i * 256 + tid
    uint32_t est[kPer];
    {
        uint4 cv[kPer][M / 16];
        float add[kPer];
#pragma unroll
        for (int i = 0; i < kPer; i++) {
            const uint32_t pos = i * kStBootThreads + tid;
            add[i] = 0.0f;
#pragma unroll
            for (int h = 0; h < M / 16; h++) cv[i][h] = make_uint4(0u, 0u, 0u, 0u);
            if (pos < ncand) {
                int r = 0;
                while (r + 1 < nprobe && s_pre[r + 1] <= pos) r++;
                const int l = s_list[r];
                const int64_t row = s_off[r] + (pos - s_pre[r]);
                const uint4* cp = reinterpret_cast<const uint4*>(codes + row * M);
#pragma unroll
                for (int h = 0; h < M / 16; h++) cv[i][h] = __ldg(cp + h);
                add[i] = st.pdis[q * nprobe + r] + __ldg(ql.sbmin + l) +
                         __ldg(ql.sbstep + l) * static_cast<float>(__ldg(ql.snorm + row)) + aq;
            }
        }
#pragma unroll
        for (int i = 0; i < kPer; i++) {
            const uint32_t pos = i * kStBootThreads + tid;
            uint32_t u = 0u;
#pragma unroll
            for (int h = 0; h < M / 16; h++) {
                const uint32_t w[4] = {cv[i][h].x, cv[i][h].y, cv[i][h].z, cv[i][h].w};
#pragma unroll
                for (int mm = 0; mm < 16; mm++)
                    u += s_lut[((w[mm >> 2] >> (8 * (mm & 3))) & 255u) * kRow + 16 * h + mm];
            }
            const float e = add[i] + static_cast<float>(u) * inv;
            uint32_t bits = __float_as_uint(fmaxf(e, 0.0f));
            if (!(e == e)) bits = 0x7f7fffffu;
            est[i] = pos < ncand ? bits : 0xffffffffu;
        }
    }
    // W rounds of a warp arg-min over the 8 x 32 estimates; the winners' positions end up in wpos (lane i holds
    // winner i; W <= 32)
    uint32_t wpos = 0xffffffffu;
#pragma unroll 1
    for (int rnd = 0; rnd < W; rnd++) {
        uint32_t best = est[0];
        int bi = 0;
#pragma unroll
        for (int i = 1; i < kPer; i++)
            if (est[i] < best) {
                best = est[i];
                bi = i;
            }
        // (estimate, lane) minimum over the warp
        uint64_t key = (static_cast<uint64_t>(best) << 32) | static_cast<uint32_t>(lane);
        for (int o = 16; o > 0; o >>= 1) {
            const uint64_t other = __shfl_xor_sync(0xffffffffu, key, o);
            key = other < key ? other : key;
        }
        const int wl = static_cast<int>(key & 31u);
        const bool valid = static_cast<uint32_t>(key >> 32) != 0xffffffffu;
        const uint32_t p = __shfl_sync(0xffffffffu, static_cast<uint32_t>(bi * kStBootThreads + tid), wl);
        if (lane == rnd) wpos = valid ? p : 0xffffffffu;
        if (lane == wl) {
#pragma unroll
            for (int i = 0; i < kPer; i++)
                if (i == bi) est[i] = 0xffffffffu;
        }
    }
    // exact distances: 16 lanes per winner (lane & 15 = sub-quantizer group), two winners per pass
    const int sub = lane & 15, half = lane >> 4;
#pragma unroll 1
    for (int pass = 0; pass < W / 2; pass++) {
        const int widx = 2 * pass + half;
        const uint32_t pos = __shfl_sync(0xffffffffu, wpos, widx);
        float tsum[M / 16];
        int l = -1;
        int64_t row = 0;
        if (pos != 0xffffffffu) {
            int r = 0;
            while (r + 1 < nprobe && s_pre[r + 1] <= pos) r++;
            l = s_list[r];
            row = s_off[r] + (pos - s_pre[r]);
        }
#pragma unroll
        for (int h = 0; h < M / 16; h++) {
            const int m = 16 * h + sub;
            float t = 0.0f;
            if (l >= 0) {
                const uint32_t cvb = __ldg(codes + row * M + m);
                const float* pc = pq + (static_cast<int64_t>(m) * 256 + cvb) * dsub;
                const float* cc = cent + static_cast<int64_t>(l) * d + m * dsub;
                for (int j = 0; j < dsub; j++) t = sqdiff_acc(t, __fsub_rn(qv[m * dsub + j], __ldg(cc + j)), __ldg(pc + j));
            }
            tsum[h] = t;
        }
        // sum over m in ascending order: m = 16 h + s, s = 0..15, through the 16 lanes of the group
        float acc = 0.0f;
#pragma unroll
        for (int h = 0; h < M / 16; h++) {
#pragma unroll
            for (int sidx = 0; sidx < 16; sidx++) {
                const float t = __shfl_sync(0xffffffffu, tsum[h], (lane & 16) | sidx);
                acc = __fadd_rn(acc, t);
            }
        }
        if (sub == 0) s_keys[wid * W + widx] = l >= 0 ? make_key(__float_as_uint(acc), pos) : kPadKey;
    }
    __syncthreads();
    if (wid == 0) {
        constexpr int R = W / 4;          // 8 W keys = 32 R
        uint64_t v[R];
#pragma unroll
        for (int r = 0; r < R; r++) v[r] = s_keys[32 * r + lane];
        TopK::warp_sort<R>(v, lane);
        // element e = 32 r + lane is v[r]: the k-th smallest is element k - 1
        const int e = k - 1;
        uint64_t kth = kPadKey;
#pragma unroll
        for (int r = 0; r < R; r++) {
            const uint64_t x = __shfl_sync(0xffffffffu, v[r], e & 31);
            if (r == (e >> 5)) kth = x;
        }
        if (lane == 0) {
            uint32_t t = kInfBits;
            uint64_t tkey = kPadKey;
            if (kth != kPadKey && static_cast<uint32_t>(kth >> 32) < kInfBits) {
                t = static_cast<uint32_t>(kth >> 32);
                tkey = kth;
            }
            qthr[q] = t;
            st.qkey[q] = tkey;
        }
    }
}

// Multi-GPU threshold exchange: thr_in[q] = the smallest bootstrap threshold any rank found for query q (all-reduce MIN
// of the distance bits).  Thresholds that came from another shard admit by distance only (tag = all ones).
__global__ void st_apply_thresholds_kernel(const uint32_t* __restrict__ thr_in, int64_t nq, uint32_t* __restrict__ qthr,
                                           uint64_t* __restrict__ qkey) {
    const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const uint32_t t = thr_in[q];
    if (t < qthr[q]) {
        qthr[q] = t;
        qkey[q] = (static_cast<uint64_t>(t) << 32) | 0xffffffffull;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// A: the filter
// ------------------------------------------------------------------------------------------------------------------
struct StCtrl {            // 64 bytes
    int work;
    int pad_[7];
    float na[kQlQ];        // -astep of the work item's queries
    float tb[kQlQ];        // threshold constants
};

template <int M>
__host__ __device__ inline size_t st_filter_smem() {
    return static_cast<size_t>(QlCfg<M>::kLutBytes) + sizeof(StCtrl) + 2 * sizeof(QlGroup);
}

// TWO: work items of at most two pairs and the 32-bit table words of ql_block16_two (M = 16)
template <int M, bool TWO = false>
__global__ void __launch_bounds__(QlCfg<M>::kT, M == 64 ? 1 : 3)
st_filter_kernel(const ScanParams p, const QlParams ql, const StParams st) {
    static_assert(!TWO || M == 16, "the two-query table layout is one 16-byte chunk per code");
    using Cfg = QlCfg<M>;
    constexpr int kT = Cfg::kT;
    constexpr int Q = kQlQ;
    extern __shared__ __align__(1024) unsigned char smem_st[];
    char* lutb = reinterpret_cast<char*>(smem_st);
    StCtrl* ctrl = reinterpret_cast<StCtrl*>(smem_st + Cfg::kLutBytes);
    QlGroup* s_grp = reinterpret_cast<QlGroup*>(ctrl + 1);

    const int tid = threadIdx.x, lane = tid & 31;
    const int r = lane & 15;
    const QlOffsets offs = TWO ? ql_make_offsets_two(lane) : ql_make_offsets(r);
    const bool x8 = (r & 8) != 0, x4 = (r & 4) != 0;
    const uint32_t bsel = (r & 3) == 0 ? 0x3210u : (r & 3) == 1 ? 0x2301u : (r & 3) == 2 ? 0x1032u : 0x0123u;
    const int ngroups = p.stats->ngroups;
    const int dsub = p.dsub;
    // this warp's open chunk of the survivor buffer (warp-uniform)
    unsigned int chunk = 0xffffffffu, fill = 0u, nrec = 0u;
    bool dead = false;

    int next_work = 0, buf = 0;
    if (tid == 0) {
        next_work = atomicAdd(&p.stats->work_counter, 1);
        if (next_work < ngroups) ql_copy_group_async(&s_grp[0], static_cast<const QlGroup*>(p.groups) + next_work);
    }
    for (;;) {
        if (tid == 0) {
            ctrl->work = next_work;
            asm volatile("cp.async.wait_all;" ::: "memory");
        }
        __syncthreads();   // also: every warp is done with the previous item's tables
        const int wk = ctrl->work;
        if (wk >= ngroups) break;
        if (tid == 0) next_work = atomicAdd(&p.stats->work_counter, 1);
        const QlGroup grp = s_grp[buf];
        int qi[Q];
        uint32_t vmask = 0u;
#pragma unroll
        for (int q = 0; q < Q; q++) {
            const bool has = grp.pair[q] >= 0;
            qi[q] = has ? grp.query[q] : grp.query[0];
            vmask |= has ? (1u << q) : 0u;
        }
        const int list = grp.list;
        const uint32_t n = grp.n;
        const uint4* lp = reinterpret_cast<const uint4*>(p.codes + grp.beg * M);
        const uint16_t* sp = ql.snorm + grp.beg;
        // first codes in flight before the set-up work
        QlCode<M> c0, c1, c2, c3;
        // TWO: the codes arrive by cp.async (L1 bypass) in a ring of kRing blocks of 256 codes that lives in the half of
        // the 256-byte table rows the two-query tables leave free.  Every thread copies, waits for and reads back ITS OWN
        // 16 bytes: no barrier, no register per byte in flight -- kRing - 1 blocks (28 KB per CTA) are outstanding, which
        // is what a DRAM-bound scan needs (one query per list: every code byte comes from HBM once)
        constexpr uint32_t kRing = 8;
        const uint32_t ring0 = static_cast<uint32_t>(__cvta_generic_to_shared(lutb + (tid >> 3) * 256 + 128 + (tid & 7) * 16));
        auto ring_issue = [&](uint32_t blk) {
            const uint32_t idx = blk * kT + tid;
            if (idx < n)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring0 + (blk & (kRing - 1)) * 8192u), "l"(lp + idx)
                             : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        auto load_sn = [&](uint32_t blk) -> uint32_t {
            const uint32_t idx = blk * kT + tid;
            return idx < n ? static_cast<uint32_t>(__ldg(sp + idx)) : 0u;
        };
        uint32_t snA = 0u, snB = 0u;
        if constexpr (TWO) {
#pragma unroll
            for (uint32_t b = 0; b + 1 < kRing; b++) ring_issue(b);
            snA = load_sn(0);
            snB = load_sn(1);
        } else {
            c0 = ql_load_code<M>(lp, sp, tid, n);
            c1 = ql_load_code<M>(lp, sp, kT + tid, n);
        }
        // pair constants: one thread per query, through shared memory (read after the barrier below)
        if (tid < Q) {
            const int q = tid;
            const QlGroup* g = &s_grp[buf];          // (dynamic index: read the shared-memory copy)
            const bool has = g->pair[q] >= 0;
            const int pr = has ? g->pair[q] : g->pair[0];
            const int qq = has ? g->query[q] : g->query[0];
            const float c_sm = __ldg(ql.sbmin + list), c_st = __ldg(ql.sbstep + list);
            const float a = __ldg(st.pdis + pr);
            const float c_am = __ldg(ql.qamin + qq), c_s = __ldg(ql.qscale + qq);
            const uint32_t thr = __ldg(p.qthr + qq);
            const float dis0 = a * (1.0f - static_cast<float>(p.d + 5) * 5.9604645e-8f);
            const float rn = sqrtf(a) * 1.00001f + ql.pmax;
            const float E = static_cast<float>(dsub + M + 8) * 5.9604645e-8f * rn * rn * 1.00001f;
            const float mag = fabsf(E) + fabsf(dis0) + fabsf(c_am) + fabsf(c_sm);
            const float base = (((E - dis0) - c_am) - c_sm) + 4.8e-7f * mag;
            ctrl->na[q] = -(c_s * c_st * 0.999999f);
            ctrl->tb[q] = has ? ql_threshold_const(thr, c_s, base, mag) : -INFINITY;
        }
        // the four queries' tables, interleaved entry-wise (see scan_qlut_kernel).  Every lane moves 16-byte pieces: one
        // LDG.128 per query fetches 8 entries (a warp reads 512 contiguous bytes of a table: 4 L1 wavefronts instead of the
        // 16 that four 32-bit loads of the same bytes take -- at 8 shards the copy was a third of the kernel's L1 traffic),
        // four STS.128 write them back interleaved.  The lanes of a quarter-warp would all write the same two 16-byte
        // columns; rotating the word order by (piece index / 2) % 4 spreads them over all eight: conflict-free.
        if constexpr (TWO) {
            // two tables -> (u_q0 | u_q1 << 16) words, written to both halves of the row's 128 bytes.  A lane's four
            // 16-byte stores go out in an order rotated by its row: the quarter-warp covers all eight columns
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const int t = i * kT + tid;
                const int row = t >> 1, piece = t & 1, rot = row & 3;
                const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(ql.qlut + static_cast<int64_t>(qi[0]) * 256 * M) + t);
                const uint4 a1 = __ldg(reinterpret_cast<const uint4*>(ql.qlut + static_cast<int64_t>(qi[1]) * 256 * M) + t);
                uint4 e0, e1;
                e0.x = __byte_perm(a0.x, a1.x, 0x5410);
                e0.y = __byte_perm(a0.x, a1.x, 0x7632);
                e0.z = __byte_perm(a0.y, a1.y, 0x5410);
                e0.w = __byte_perm(a0.y, a1.y, 0x7632);
                e1.x = __byte_perm(a0.z, a1.z, 0x5410);
                e1.y = __byte_perm(a0.z, a1.z, 0x7632);
                e1.z = __byte_perm(a0.w, a1.w, 0x5410);
                e1.w = __byte_perm(a0.w, a1.w, 0x7632);
                char* dst = lutb + row * 256 + piece * 32;
#pragma unroll
                for (int s2 = 0; s2 < 4; s2++) {
                    const int sp2 = (s2 + rot) & 3, cp = sp2 >> 1, jj = sp2 & 1;
                    *reinterpret_cast<uint4*>(dst + cp * 64 + jj * 16) = jj ? e1 : e0;
                }
            }
        } else
        {
            constexpr int kPieces = M / 8;            // 16-byte pieces per table row = pieces per thread
            constexpr int kBatch = 2;
#pragma unroll 1
            for (int i0 = 0; i0 < kPieces; i0 += kBatch) {
                uint4 a[kBatch][Q];
#pragma unroll
                for (int i = 0; i < kBatch; i++) {
                    const int t = (i0 + i) * kT + tid;
#pragma unroll
                    for (int q = 0; q < Q; q++)
                        a[i][q] = __ldg(reinterpret_cast<const uint4*>(ql.qlut + static_cast<int64_t>(qi[q]) * 256 * M) + t);
                }
#pragma unroll
                for (int i = 0; i < kBatch; i++) {
                    const int t = (i0 + i) * kT + tid;
                    const int row = t / kPieces, piece = t % kPieces, h = piece >> 1, half = piece & 1;
                    const int rot = (t >> 1) & 3;
#pragma unroll
                    for (int q = 0; q < Q; q++) {
                        uint4 v = a[i][q];
                        if (rot & 1) v = make_uint4(v.y, v.z, v.w, v.x);
                        if (rot & 2) v = make_uint4(v.z, v.w, v.x, v.y);
                        a[i][q] = v;
                    }
                    char* dst = lutb + (h >> 1) * kQlPlaneBytes + row * 256 + (h & 1) * 128 + half * 64;
                    const uint32_t w[4][Q] = {{a[i][0].x, a[i][1].x, a[i][2].x, a[i][3].x},
                                              {a[i][0].y, a[i][1].y, a[i][2].y, a[i][3].y},
                                              {a[i][0].z, a[i][1].z, a[i][2].z, a[i][3].z},
                                              {a[i][0].w, a[i][1].w, a[i][2].w, a[i][3].w}};
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        uint4 e;
                        e.x = __byte_perm(w[j][0], w[j][1], 0x5410);
                        e.y = __byte_perm(w[j][2], w[j][3], 0x5410);
                        e.z = __byte_perm(w[j][0], w[j][1], 0x7632);
                        e.w = __byte_perm(w[j][2], w[j][3], 0x7632);
                        *reinterpret_cast<uint4*>(dst + ((j + rot) & 3) * 16) = e;
                    }
                }
            }
        }
        __syncthreads();
        buf ^= 1;
        if (tid == 0 && next_work < ngroups)
            ql_copy_group_async(&s_grp[buf], static_cast<const QlGroup*>(p.groups) + next_work);
        float na[Q], tb[Q];
#pragma unroll
        for (int q = 0; q < Q; q++) {
            na[q] = ctrl->na[q];
            tb[q] = ctrl->tb[q];
        }

        auto test = [&](const QlCode<M>& c) -> uint32_t {
            if constexpr (TWO) {
                const uint32_t lb = ql_block16_two(lutb, c.v[0], x8, x4, bsel, offs);
                const float f0 = __uint_as_float(__byte_perm(lb, 0x4b000000u, 0x7410));
                const float f1 = __uint_as_float(__byte_perm(lb, 0x4b000000u, 0x7432));
                const float vs = static_cast<float>(c.s);
                const bool h0 = !(f0 > fmaf(vs, na[0], tb[0])), h1 = !(f1 > fmaf(vs, na[1], tb[1]));
                return (h0 ? 1u : 0u) | (h1 ? 2u : 0u);
            }
            float f[4];
            ql_bounds_as_floats<M>(lutb, c.v, x8, x4, bsel, offs, f);
            const float vs = static_cast<float>(c.s);
            const bool h0 = !(f[0] > fmaf(vs, na[0], tb[0])), h1 = !(f[1] > fmaf(vs, na[1], tb[1])),
                       h2 = !(f[2] > fmaf(vs, na[2], tb[2])), h3 = !(f[3] > fmaf(vs, na[3], tb[3]));
            return (h0 ? 1u : 0u) | (h1 ? 2u : 0u) | (h2 ? 4u : 0u) | (h3 ? 8u : 0u);
        };
        // survivors go straight to this warp's chunk of the global record buffer
        auto append = [&](bool hit, uint32_t entry) {
            const unsigned bal = __ballot_sync(0xffffffffu, hit);
            if (bal == 0u || dead) return;
            const unsigned np = __popc(bal);
            if (chunk == 0xffffffffu || fill + np > kStChunk) {
                unsigned int c = 0u;
                if (lane == 0) {
                    if (chunk != 0xffffffffu) st.sfill[chunk] = fill;
                    c = atomicAdd(&st.ctr->nchunks, 1u);
                }
                c = __shfl_sync(0xffffffffu, c, 0);
                if (c >= st.max_chunks) {
                    if (lane == 0) atomicOr(&st.ctr->overflow, 2);
                    dead = true;
                    chunk = 0xffffffffu;
                    return;
                }
                chunk = c;
                fill = 0u;
            }
            if (hit) st.srec[static_cast<size_t>(chunk) * kStChunk + fill + __popc(bal & lanemask_lt())] =
                         make_uint2(static_cast<uint32_t>(wk), entry);
            fill += np;
            nrec += np;
        };

        const uint32_t nblk = (n + kT - 1) / kT;
        if constexpr (TWO) {
#define ST_ITER2(SN, B)                                                                              \
    {                                                                                                \
        asm volatile("cp.async.wait_group 6;" ::: "memory");                                         \
        QlCode<M> c;                                                                                 \
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"                                      \
                     : "=r"(c.v[0].x), "=r"(c.v[0].y), "=r"(c.v[0].z), "=r"(c.v[0].w)                \
                     : "r"(ring0 + ((B) & (kRing - 1)) * 8192u)                                      \
                     : "memory");                                                                    \
        c.s = SN;                                                                                    \
        ring_issue((B) + kRing - 1);                                                                 \
        SN = load_sn((B) + 2);                                                                       \
        const uint32_t idx = (B) * kT + tid;                                                         \
        const uint32_t m_ = test(c);                                                                 \
        append(idx < n && m_ != 0u, (idx << 4) | m_);                                                \
    }
            static_assert(kRing == 8, "cp.async.wait_group 6 = kRing - 2");
#pragma unroll 1
            for (uint32_t b = 0; b < nblk; b += 2) {
                ST_ITER2(snA, b)
                if (b + 1 < nblk) ST_ITER2(snB, b + 1)
            }
#undef ST_ITER2
        } else {
#define ST_ITER(CUR, LOADTO, TB)                                                                     \
    {                                                                                                \
        LOADTO = ql_load_code<M>(lp, sp, base + (TB + 2) * kT, n);                                   \
        const uint32_t idx = base + TB * kT;                                                         \
        const uint32_t m_ = test(CUR);                                                               \
        append(idx < n && m_ != 0u, (idx << 4) | m_);                                                \
    }
#pragma unroll 1
        for (uint32_t t0 = 0; t0 < nblk; t0 += 4) {
            const uint32_t base = t0 * kT + tid;
            ST_ITER(c0, c2, 0)
            if (t0 + 1 < nblk) ST_ITER(c1, c3, 1)
            if (t0 + 2 < nblk) ST_ITER(c2, c0, 2)
            if (t0 + 3 < nblk) ST_ITER(c3, c1, 3)
        }
#undef ST_ITER
        }
    }
    if (lane == 0) {
        if (chunk != 0xffffffffu) st.sfill[chunk] = fill;
        if (nrec) atomicAdd(&st.ctr->records, static_cast<unsigned long long>(nrec));
    }
}

// ------------------------------------------------------------------------------------------------------------------
// A2: the filter for lists probed by ONE or TWO queries (few queries per list: small batches x nprobe, the RAG serving
// regime, where every code byte comes from DRAM once).  M = 16.
//   * Table entry = one 32-bit word (u_q0, u_q1): ONE LDS.32 wavefront serves two queries x 32 codes (the four-query
//     layout spends an LDS.64 -- two wavefronts -- on the same look-up and wastes half of them on empty query slots).
//     Row = code value, 256-byte stride, word (lane / 16) * 16 + (r ^ step): 32 distinct banks for any code bytes.
//   * Code tiles arrive by BULK ASYNC COPY (cp.async.bulk global -> shared, completion on an mbarrier): a ring of kF2Stages
//     tiles of 256 codes per CTA, issued by one thread, so kF2Stages x 4 KB per CTA are in flight without holding a
//     register and without passing through the LSU on the way in.  Every warp releases a stage with one arrive on its
//     `empty` mbarrier after it has copied its codes to registers (LDS.128, conflict-free).
// Everything else (thresholds, survivor records, exact evaluation downstream) is st_filter_kernel's.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kF2Stages = 6;
constexpr int kF2Tile = 256;                       // codes per stage (= threads)
constexpr int kF2TileBytes = kF2Tile * 16;

__device__ __forceinline__ uint32_t st_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void st_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void st_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void st_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded wait: a protocol error must end in a trap, not in a hung GPU
__device__ __forceinline__ void st_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; spin++) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (spin > (1u << 28)) __trap();
    }
}
__device__ __forceinline__ void st_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

struct StF2Ctrl {
    int work;
    int pad_[3];
    float na[2];
    float tb[2];
    unsigned long long full[kF2Stages];
    unsigned long long empty[kF2Stages];
};

__host__ __device__ inline size_t st_filter2_smem() {
    return static_cast<size_t>(kQlPlaneBytes) + kF2Stages * kF2TileBytes + sizeof(StF2Ctrl) + 2 * sizeof(QlGroup);
}

// offsets of the 16 look-up steps for the 32-bit table: (lane / 16) * 64 + (r ^ p) * 4 (see ql_make_offsets)
__device__ __forceinline__ QlOffsets st_make_offsets32(int lane) {
    const int r = lane & 15, h = lane >> 4;
    QlOffsets f;
#pragma unroll
    for (int i = 0; i < 6; i++) {
        uint32_t v = 0u;
#pragma unroll
        for (int b = 0; b < 3; b++) {
            const int p = 3 * i + b;
            if (p < 16) v |= static_cast<uint32_t>(h * 64 + (r ^ p) * 4) << (8 * b);
        }
        f.o[i] = v;
    }
    return f;
}
template <int B, int P>
__device__ __forceinline__ uint32_t st_lookup32(const char* __restrict__ lutb, uint32_t w, const QlOffsets& f) {
    const uint32_t a = __byte_perm(w, f.o[P / 3], 0x7700u | (B << 4) | (4 + P % 3));
    return *reinterpret_cast<const uint32_t*>(lutb + a);
}
__device__ __forceinline__ uint32_t st_block16_32(const char* __restrict__ lutb, const uint4& code, bool x8, bool x4,
                                                  uint32_t bsel, const QlOffsets& f) {
    const uint32_t y0 = x8 ? code.z : code.x, y1 = x8 ? code.w : code.y, y2 = x8 ? code.x : code.z,
                   y3 = x8 ? code.y : code.w;
    const uint32_t z0 = x4 ? y1 : y0, z1 = x4 ? y0 : y1, z2 = x4 ? y3 : y2, z3 = x4 ? y2 : y3;
    const uint32_t w0 = __byte_perm(z0, 0u, bsel), w1 = __byte_perm(z1, 0u, bsel), w2 = __byte_perm(z2, 0u, bsel),
                   w3 = __byte_perm(z3, 0u, bsel);
    uint32_t s = 0u;
#define ST_STEP(W, B, P) s += st_lookup32<B, P>(lutb, W, f);
    ST_STEP(w0, 0, 0) ST_STEP(w0, 1, 1) ST_STEP(w0, 2, 2) ST_STEP(w0, 3, 3)
    ST_STEP(w1, 0, 4) ST_STEP(w1, 1, 5) ST_STEP(w1, 2, 6) ST_STEP(w1, 3, 7)
    ST_STEP(w2, 0, 8) ST_STEP(w2, 1, 9) ST_STEP(w2, 2, 10) ST_STEP(w2, 3, 11)
    ST_STEP(w3, 0, 12) ST_STEP(w3, 1, 13) ST_STEP(w3, 2, 14) ST_STEP(w3, 3, 15)
#undef ST_STEP
    return s;
}

__global__ void __launch_bounds__(256, 2)
st_filter2_kernel(const ScanParams p, const QlParams ql, const StParams st) {
    constexpr int M = 16, kT = 256;
    extern __shared__ __align__(1024) unsigned char smem_st[];
    char* lutb = reinterpret_cast<char*>(smem_st);
    unsigned char* ring = smem_st + kQlPlaneBytes;                               // kF2Stages tiles of 4 KB
    StF2Ctrl* ctrl = reinterpret_cast<StF2Ctrl*>(ring + kF2Stages * kF2TileBytes);
    QlGroup* s_grp = reinterpret_cast<QlGroup*>(ctrl + 1);

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int r = lane & 15;
    const QlOffsets offs = st_make_offsets32(lane);
    const bool x8 = (r & 8) != 0, x4 = (r & 4) != 0;
    const uint32_t bsel = (r & 3) == 0 ? 0x3210u : (r & 3) == 1 ? 0x2301u : (r & 3) == 2 ? 0x1032u : 0x0123u;
    const int ngroups = p.stats->ngroups;
    const int dsub = p.dsub;
    unsigned int chunk = 0xffffffffu, fill = 0u, nrec = 0u;
    bool dead = false;
    uint32_t g = 0;          // tiles this CTA has consumed so far: stage = g % kF2Stages, phase = (g / kF2Stages) & 1
    uint32_t gi = 0;         // tiles issued so far (thread 0)

    if (tid == 0) {
        for (int s = 0; s < kF2Stages; s++) {
            st_mbar_init(st_smem_u32(&ctrl->full[s]), 1);
            st_mbar_init(st_smem_u32(&ctrl->empty[s]), kT / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    int next_work = 0, buf = 0;
    if (tid == 0) {
        next_work = atomicAdd(&p.stats->work_counter, 1);
        if (next_work < ngroups) ql_copy_group_async(&s_grp[0], static_cast<const QlGroup*>(p.groups) + next_work);
    }
    for (;;) {
        if (tid == 0) {
            ctrl->work = next_work;
            asm volatile("cp.async.wait_all;" ::: "memory");
        }
        __syncthreads();   // every warp is done with the previous item's tables; all its tiles are consumed
        const int wk = ctrl->work;
        if (wk >= ngroups) break;
        if (tid == 0) next_work = atomicAdd(&p.stats->work_counter, 1);
        const QlGroup grp = s_grp[buf];
        const bool has1 = grp.pair[1] >= 0;
        const int q0 = grp.query[0], q1 = has1 ? grp.query[1] : grp.query[0];
        const int list = grp.list;
        const uint32_t n = grp.n;
        const uint32_t nblk = (n + kT - 1) / kT;
        const unsigned char* lcodes = p.codes + grp.beg * M;
        const uint16_t* sp = ql.snorm + grp.beg;
        // producer: the first tiles of this list go out before the table work starts
        auto issue = [&](uint32_t blk) {   // thread 0 only
            const uint32_t s = gi % kF2Stages;
            if (gi >= kF2Stages) st_mbar_wait(st_smem_u32(&ctrl->empty[s]), ((gi / kF2Stages) - 1) & 1);
            const uint32_t cnt = min(static_cast<uint32_t>(kT), n - blk * kT);
            st_mbar_expect_tx(st_smem_u32(&ctrl->full[s]), cnt * 16u);
            st_bulk_load(st_smem_u32(ring + s * kF2TileBytes), lcodes + static_cast<size_t>(blk) * kF2TileBytes, cnt * 16u,
                         st_smem_u32(&ctrl->full[s]));
            gi++;
        };
        uint32_t issued = 0;   // tiles of THIS list issued (thread 0)
        if (tid == 0)
            for (; issued < nblk && issued < kF2Stages - 1; issued++) issue(issued);
        // pair constants: one thread per query
        if (tid < 2) {
            const int q = tid;
            const bool has = q == 0 || has1;
            const int pr = has ? grp.pair[0] * (1 - q) + (q ? grp.pair[1] : 0) : grp.pair[0];
            const int qq = q == 0 ? q0 : q1;
            const float c_sm = __ldg(ql.sbmin + list), c_st = __ldg(ql.sbstep + list);
            const float a = __ldg(st.pdis + pr);
            const float c_am = __ldg(ql.qamin + qq), c_s = __ldg(ql.qscale + qq);
            const uint32_t thr = __ldg(p.qthr + qq);
            const float dis0 = a * (1.0f - static_cast<float>(p.d + 5) * 5.9604645e-8f);
            const float rn = sqrtf(a) * 1.00001f + ql.pmax;
            const float E = static_cast<float>(dsub + M + 8) * 5.9604645e-8f * rn * rn * 1.00001f;
            const float mag = fabsf(E) + fabsf(dis0) + fabsf(c_am) + fabsf(c_sm);
            const float base = (((E - dis0) - c_am) - c_sm) + 4.8e-7f * mag;
            ctrl->na[q] = -(c_s * c_st * 0.999999f);
            ctrl->tb[q] = has ? ql_threshold_const(thr, c_s, base, mag) : -INFINITY;
        }
        // the two queries' tables: word (c, h, slot m) = (u_q0[c][m], u_q1[c][m]); a lane handles one row's slot pair
        // (2 sp2, 2 sp2 + 1) and writes it to both half-warp copies
        {
            const int sp2 = lane & 7;
#pragma unroll 1
            for (int i0 = 0; i0 < 8; i0 += 4) {
                uint32_t a0[4], a1[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int it = wid + (i0 + i) * (kT / 32);
                    const int row = (it << 2) + (lane >> 3);
                    a0[i] = __ldg(reinterpret_cast<const uint32_t*>(ql.qlut + (static_cast<int64_t>(q0) * 256 + row) * M) + sp2);
                    a1[i] = __ldg(reinterpret_cast<const uint32_t*>(ql.qlut + (static_cast<int64_t>(q1) * 256 + row) * M) + sp2);
                }
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int it = wid + (i0 + i) * (kT / 32);
                    const int row = (it << 2) + (lane >> 3);
                    const uint2 e = make_uint2(__byte_perm(a0[i], a1[i], 0x5410), __byte_perm(a0[i], a1[i], 0x7632));
                    *reinterpret_cast<uint2*>(lutb + row * 256 + sp2 * 8) = e;
                    *reinterpret_cast<uint2*>(lutb + row * 256 + 64 + sp2 * 8) = e;
                }
            }
        }
        __syncthreads();
        buf ^= 1;
        if (tid == 0 && next_work < ngroups)
            ql_copy_group_async(&s_grp[buf], static_cast<const QlGroup*>(p.groups) + next_work);
        const float na0 = ctrl->na[0], na1 = ctrl->na[1], tb0 = ctrl->tb[0], tb1 = ctrl->tb[1];

        uint32_t sn0 = tid < n ? __ldg(sp + tid) : 0u, sn1 = kT + tid < n ? __ldg(sp + kT + tid) : 0u;
        // the code of block b + 1 is copied out of the ring (and its stage released) while block b is looked up
        auto fetch = [&](uint32_t blk) -> uint4 {
            const uint32_t s = g % kF2Stages;
            st_mbar_wait(st_smem_u32(&ctrl->full[s]), (g / kF2Stages) & 1);
            uint4 c = make_uint4(0u, 0u, 0u, 0u);
            if (blk * kT + tid < n) c = *reinterpret_cast<const uint4*>(ring + s * kF2TileBytes + tid * 16);
            g++;
            return c;
        };
        uint4 code = fetch(0);
        uint32_t rel = (g - 1) % kF2Stages;     // stage to release once `code` has been consumed
#pragma unroll 1
        for (uint32_t blk = 0; blk < nblk; blk++) {
            const uint32_t idx = blk * kT + tid;
            uint4 code_next = make_uint4(0u, 0u, 0u, 0u);
            uint32_t rel_next = 0u;
            if (blk + 1 < nblk) {
                code_next = fetch(blk + 1);
                rel_next = (g - 1) % kF2Stages;
            }
            const uint32_t sn = sn0;
            sn0 = sn1;
            sn1 = idx + 2 * kT < n ? __ldg(sp + idx + 2 * kT) : 0u;
            const uint32_t lb = st_block16_32(lutb, code, x8, x4, bsel, offs);
            // the look-ups consumed the code registers, so this warp's reads of that stage are complete: release it,
            // and (thread 0) refill the ring
            __syncwarp();
            if (lane == 0) st_mbar_arrive(st_smem_u32(&ctrl->empty[rel]));
            if (tid == 0 && issued < nblk) {
                issue(issued);
                issued++;
            }
            code = code_next;
            rel = rel_next;
            const float vs = static_cast<float>(sn);
            const float f0 = __uint_as_float(__byte_perm(lb, 0x4b000000u, 0x7410)),
                        f1 = __uint_as_float(__byte_perm(lb, 0x4b000000u, 0x7432));
            const bool h0 = !(f0 > fmaf(vs, na0, tb0)), h1 = !(f1 > fmaf(vs, na1, tb1));
            const bool hit = idx < n && (h0 || h1);
            const unsigned bal = __ballot_sync(0xffffffffu, hit);
            if (bal != 0u && !dead) {
                const unsigned np = __popc(bal);
                if (chunk == 0xffffffffu || fill + np > kStChunk) {
                    unsigned int c = 0u;
                    if (lane == 0) {
                        if (chunk != 0xffffffffu) st.sfill[chunk] = fill;
                        c = atomicAdd(&st.ctr->nchunks, 1u);
                    }
                    c = __shfl_sync(0xffffffffu, c, 0);
                    if (c >= st.max_chunks) {
                        if (lane == 0) atomicOr(&st.ctr->overflow, 2);
                        dead = true;
                        chunk = 0xffffffffu;
                    } else {
                        chunk = c;
                        fill = 0u;
                    }
                }
                if (!dead) {
                    if (hit) st.srec[static_cast<size_t>(chunk) * kStChunk + fill + __popc(bal & lanemask_lt())] =
                                 make_uint2(static_cast<uint32_t>(wk), (idx << 4) | (h0 ? 1u : 0u) | (h1 ? 2u : 0u));
                    fill += np;
                    nrec += np;
                }
            }
        }
    }
    if (lane == 0) {
        if (chunk != 0xffffffffu) st.sfill[chunk] = fill;
        if (nrec) atomicAdd(&st.ctr->records, static_cast<unsigned long long>(nrec));
    }
}

// ------------------------------------------------------------------------------------------------------------------
// B: exact evaluation of the survivors.  One thread per record.  The 16 (M) codebook rows a record needs are random
// gathers: from global memory every lane of an LDG touches its own cache line (32 L1 wavefronts per instruction; ncu on
// C2: that is the kernel's whole run time), so the PQ codebook is staged in shared memory when it fits (128 KB at
// d = 128, rows padded so that random rows spread over the banks) and the gathers become LDS.
// ------------------------------------------------------------------------------------------------------------------
__host__ __device__ inline int st_eval_row_stride(int dsub) { return dsub + ((dsub & 3) == 0 ? 4 : (dsub & 1) == 0 ? 2 : 1); }
__host__ __device__ inline size_t st_eval_smem(int M, int dsub) {
    return sizeof(float) * static_cast<size_t>(M) * 256 * st_eval_row_stride(dsub);
}
constexpr int kStEvalThreads = 1024;
constexpr size_t kStEvalSmemMax = 200 * 1024;

// pq_rows: codebook rows with `stride` floats between them (global: stride = dsub; shared: padded)
template <int M>
__device__ __forceinline__ float st_exact_rows(const uint8_t* __restrict__ code, const float* __restrict__ q,
                                               const float* __restrict__ c, const float* __restrict__ pq_rows, int dsub,
                                               int stride) {
    float acc = 0.0f;
#pragma unroll 1
    for (int m0 = 0; m0 < M; m0 += 4) {
        const uint32_t cw = __ldg(reinterpret_cast<const uint32_t*>(code + m0));
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int m = m0 + b;
            const uint32_t cv = (cw >> (8 * b)) & 255u;
            const float* pc = pq_rows + (static_cast<int64_t>(m) * 256 + cv) * stride;
            const float* qq = q + m * dsub;
            const float* cc = c + m * dsub;
            float t = 0.0f;
            if ((dsub & 3) == 0) {
                for (int j = 0; j < dsub; j += 4) {
                    const float4 pv = *reinterpret_cast<const float4*>(pc + j);
                    const float4 qv = __ldg(reinterpret_cast<const float4*>(qq + j));
                    const float4 cv4 = __ldg(reinterpret_cast<const float4*>(cc + j));
                    t = sqdiff_acc(t, __fsub_rn(qv.x, cv4.x), pv.x);
                    t = sqdiff_acc(t, __fsub_rn(qv.y, cv4.y), pv.y);
                    t = sqdiff_acc(t, __fsub_rn(qv.z, cv4.z), pv.z);
                    t = sqdiff_acc(t, __fsub_rn(qv.w, cv4.w), pv.w);
                }
            } else if ((dsub & 1) == 0) {
                for (int j = 0; j < dsub; j += 2) {
                    const float2 pv = *reinterpret_cast<const float2*>(pc + j);
                    const float2 qv = __ldg(reinterpret_cast<const float2*>(qq + j));
                    const float2 cv2 = __ldg(reinterpret_cast<const float2*>(cc + j));
                    t = sqdiff_acc(t, __fsub_rn(qv.x, cv2.x), pv.x);
                    t = sqdiff_acc(t, __fsub_rn(qv.y, cv2.y), pv.y);
                }
            } else {
                for (int j = 0; j < dsub; j++) t = sqdiff_acc(t, __fsub_rn(__ldg(qq + j), __ldg(cc + j)), pc[j]);
            }
            acc = __fadd_rn(acc, t);
        }
    }
    return acc;
}

template <int M, bool SMEM_PQ>
__global__ void __launch_bounds__(kStEvalThreads)
st_eval_kernel(const ScanParams p, const StParams st) {
    extern __shared__ __align__(16) float st_spq[];
    const int tid = threadIdx.x;
    const int stride = SMEM_PQ ? st_eval_row_stride(p.dsub) : p.dsub;
    const float* pq_rows = p.pq;
    if (SMEM_PQ) {
        if ((p.dsub & 3) == 0) {
            const int v4 = p.dsub >> 2;
            for (int e = tid; e < M * 256 * v4; e += kStEvalThreads)
                *reinterpret_cast<float4*>(st_spq + (e / v4) * stride + 4 * (e % v4)) = __ldg(reinterpret_cast<const float4*>(p.pq) + e);
        } else {
            for (int e = tid; e < M * 256 * p.dsub; e += kStEvalThreads) st_spq[(e / p.dsub) * stride + e % p.dsub] = __ldg(p.pq + e);
        }
        pq_rows = st_spq;
        __syncthreads();
    }
    const unsigned int nchunks = min(st.ctr->nchunks, st.max_chunks);
    constexpr unsigned int kPer = kStEvalThreads / kStChunk;   // chunks per CTA and iteration
    unsigned long long nev = 0ull;
    for (unsigned int c0 = blockIdx.x * kPer; c0 < nchunks; c0 += gridDim.x * kPer) {
        const unsigned int c = c0 + (tid >> 6);
        const unsigned int slot = tid & 63;
        if (c >= nchunks || slot >= st.sfill[c]) continue;
        const uint2 rec = st.srec[static_cast<size_t>(c) * kStChunk + slot];
        const QlGroup* g = static_cast<const QlGroup*>(p.groups) + rec.x;
        const uint32_t idx = rec.y >> 4;
        uint32_t bits = rec.y & 15u;
        const int list = g->list;
        const int64_t row = g->beg + idx;
        const uint8_t* code = p.codes + row * M;
        const float* crow = p.cent + static_cast<int64_t>(list) * p.d;
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1u;
            const int pair = g->pair[b], q = g->query[b];
            const uint32_t db = __float_as_uint(
                st_exact_rows<M>(code, p.xq + static_cast<int64_t>(q) * p.d, crow, pq_rows, p.dsub, stride));
            nev++;
            const uint64_t key = make_key(db, __ldg(st.prefix + pair) + idx);
            if (key <= __ldg(st.qkey + q)) {
                const unsigned int s = atomicAdd(&st.qcnt[q], 1u);
                if (s < static_cast<unsigned int>(st.capq)) {
                    st.slab[static_cast<size_t>(q) * st.capq + s] = key;
                } else if (s == static_cast<unsigned int>(st.capq)) {
                    st.qflag[q] = 1;
                    atomicOr(&st.ctr->overflow, 4);
                }
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) nev += __shfl_xor_sync(0xffffffffu, nev, o);
    if ((tid & 31) == 0 && nev) atomicAdd(&st.ctr->evals, nev);
}

// ------------------------------------------------------------------------------------------------------------------
// C: per-query selection + id lookup
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kStSelThreads)
st_select_kernel(const StParams st, const int32_t* __restrict__ probe, const int64_t* __restrict__ offsets, int nprobe,
                 int k) {
    if (threadIdx.x == 0) {
        const unsigned int c = st.qcnt[blockIdx.x];
        atomicAdd(&st.ctr->kept, static_cast<unsigned long long>(c));
        atomicMax(&st.ctr->maxkept, c);
        if (st.qflag[blockIdx.x] != 0) atomicAdd(&st.ctr->nflag, 1u);
    }
    // the fallback launches answer: everything after a batch-wide overflow, this query after a slab overflow
    if ((*reinterpret_cast<const volatile int*>(&st.ctr->overflow) & 3) != 0 || st.qflag[blockIdx.x] != 0) return;
    extern __shared__ __align__(16) unsigned char smem_sel[];
    TopK tk;
    tk.bind(smem_sel, k, kStSelCap);
    const int tid = threadIdx.x;
    const int64_t q = blockIdx.x;
    if (tid == 0) tk.reset(kInfBits);
    __syncthreads();
    const unsigned int n = min(st.qcnt[q], static_cast<unsigned int>(st.capq));
    const uint64_t* keys = st.slab + static_cast<size_t>(q) * st.capq;
    uint32_t thr = kInfBits;
    for (unsigned int base = 0; base < n; base += kStSelThreads * 4) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const unsigned int i = base + u * kStSelThreads + tid;
            uint64_t key = kPadKey;
            if (i < n) key = keys[i];
            tk.push(i < n && static_cast<uint32_t>(key >> 32) <= thr, key);
        }
        tk.sync_and_flush_if_over<kStSelThreads>(kStSelCap - kStSelThreads * 4, kInfBits);
        thr = tk.threshold();
    }
    __syncthreads();
    tk.flush<kStSelThreads>(kInfBits);
    const int nb = tk.count();
    const uint64_t* s = tk.sorted();
    const uint32_t* pre = st.prefix + q * nprobe;
    for (int i = tid; i < k; i += kStSelThreads) {
        float dv = FLT_MAX;
        int64_t id = -1;
        if (i < nb) {
            const uint32_t tag = static_cast<uint32_t>(s[i] & 0xffffffffu);
            int lo = 0, hi = nprobe - 1;      // last rank whose first scan position is <= tag
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (pre[mid] <= tag) lo = mid;
                else hi = mid - 1;
            }
            const int64_t pos = offsets[probe[q * nprobe + lo]] + (tag - pre[lo]);
            id = st.ids ? st.ids[pos] : pos;
            dv = __uint_as_float(static_cast<uint32_t>(s[i] >> 32));
        }
        st.D[q * k + i] = dv;
        st.I[q * k + i] = id;
    }
}

}  // namespace b200
