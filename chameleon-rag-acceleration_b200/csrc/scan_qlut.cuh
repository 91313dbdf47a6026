// scan_qlut.cuh -- ADC scan with a PER-QUERY integer lower-bound filter and exact evaluation of the survivors.
//
// Why.  Round 1's filter kernel (scan_quad.cuh) builds a table per (query, list) pair: 3*256*d flops, the whole PQ
// codebook (256*d*4 bytes) re-read from L2 and a 64 KB exact table written to a global scratch, for every work
// item.  Sharding the lists by vector (the reference's co.shard = True, bench_gpu_performance_OSDI.py:586-604) divides
// the codes a GPU scans but not the pairs, so that fixed cost is what every shard repeats -- 8 GPUs gave 2.7x.
//
// How.  The distance the oracle evaluates, ||q - c - p||^2 with p = (pq[m][code_m])_m, splits in real arithmetic into
//
//     ||q - c||^2  +  [ ||p||^2 + 2 (c - mu).p ]  +  sum_m [ -2 (q - mu)_m . pq[m][code_m] ]
//        dis0(q, list)        SB (stored vector)             A_q[m][code_m]   (query only)
//
// (mu = mean of the coarse centroids, any fixed vector works).  Nothing in it depends on the PAIR except one scalar:
//   * A_q is built ONCE per query and batch (ql_query_tables_kernel), quantised to 11 bits (M = 16, 32,
//     64):  u[m][c] = floor((A[m][c] + B_m) * s_q), B_m >= |A[m][.]| (Cauchy-Schwarz);  a work item only copies the
//     tables of its queries into shared memory (8 KB per query at M = 16, from L2);
//   * SB is built ONCE per stored vector when the lists are installed (ql_sb_build_kernel, float64), kept as 16 bits
//     on a per-list grid, rounded DOWN:  SB >= sbmin[l] + sbstep[l] * v;
//   * a code can only be among the results if
//         sum_m u[m][code_m]  <=  s_q * (thr + E - dis0 + sum_m B_m - sbmin - sbstep * v)
//     (thr: the query's current k-th best distance; E: rounding slack, below).  The left side is the same
//     conflict-free packed-integer look-up sum as scan_quad.cuh (one LDS.64 serves four queries); the right side is
//     one FFMA per (code, query), compared in the "magic number" float domain (2^23 + integer) so that no
//     conversion or clamp is needed;
//   * survivors (about 1 % of the codes once a threshold exists) are evaluated EXACTLY and on the fly from the
//     codebook: T[m] = sum_j ((q - c)[m*dsub+j] - pq[m][code_m][j])^2, j ascending, then sum_m ascending, every
//     operation rounded separately -- bit-identical to the oracle (oracle/ivfpq_oracle.c; reference
//     LUT_construction.hpp:180-209, ADC.hpp:75-99) -- and go through the same top-k machinery (strict <,
//     priority_queue_L1.hpp:65-75).  No exact table exists any more.
// The result set is exactly the oracle's; only the amount of exact work depends on the filter.
//
// Rounding slack (everything that is computed in floating point is pushed in the conservative direction):
//   * oracle distance D_o vs the real value D*:  D_o >= D* - E,  E = (dsub + M + 8) 2^-24 (||r|| + Pmax)^2, r = fl(q - c),
//     Pmax >= ||p|| for every code word (errors of fl(q - c) - p, of the squares and of the two nested sequential sums);
//   * dis0 is summed from the same r: real ||q - c||^2 >= dis0 (1 - (d + 5) 2^-24);
//   * A is evaluated with fp32 FMAs: |A_fl - A_real| <= (dsub + 3) 2^-23 ||(q - mu)_m|| ||pq[m][c]||, summed over m and
//     taken off the sum of the offsets once per query; the quantiser multiplies by (1 - 2^-20) before the floor;
//   * SB is evaluated in float64 from the fp32 inputs (products exact, sums to 2^-53) and the grid point is checked
//     against it in float64 before it is stored;
//   * the pair constants are rounded up by 2^-21 of the magnitudes involved, the threshold by another +2 units.
// tests/test_qlut_bound.py restates all of this in numpy and checks the inequality on adversarial data.
#pragma once
#include "scan_types.cuh"

namespace b200 {

constexpr int kQlQ = 4;                           // queries per work item
constexpr int kQlPlaneBytes = 256 * 256;          // one table plane: 256 code values x 256-byte rows
constexpr uint32_t kQlMaxList = 1u << 28;         // survivor entry = (offset << 4) | query mask
constexpr int kQlSurvCap = 2304;                  // one tile of 2048 codes + the drain trigger
constexpr float kQlMagic = 8388608.0f;            // 2^23

template <int M>
struct QlCfg {
    static_assert(M == 16 || M == 32 || M == 64, "M = 16, 32 or 64");
    static constexpr int kT = 256;                          // threads per CTA
    static constexpr int kChunks = M / 16;                  // 16-byte chunks of a code; one 128-byte table row each
    static constexpr int kPlanes = (kChunks + 1) / 2;       // two chunk rows per 256-byte plane row
    static constexpr int kLutBytes = kPlanes * kQlPlaneBytes;
    static constexpr uint32_t kQMax = 2047u;                // 11-bit entries: the 16 of a chunk sum below 2^15 in a packed
                                                            // half-word; chunks are widened to 32 bits before they are added
    static constexpr int kTileBlocks = 1024 / kT;           // blocks of kT codes between survivor checks
};

// work item: up to four (query, probe) pairs of the same list
struct __align__(16) QlGroup {
    int pair[kQlQ];   // -1: unused slot (always at the end)
    int query[kQlQ];  // pair / nprobe
    int list;
    uint32_t n;       // list length (> 0)
    int64_t beg;      // first row of the list in codes / ids / snorm
};
static_assert(sizeof(QlGroup) == 48, "QlGroup is three 16-byte words");

struct QlParams {
    // per index
    const uint16_t* snorm;    // (ntotal) per-vector term on the per-list grid
    const float* sbmin;       // (nlist)
    const float* sbstep;      // (nlist)
    float pmax;               // >= ||p|| for every code word
    // per batch
    const uint16_t* qlut;     // (nq, 256, M) quantised per-query tables, row = code value
    const float* qscale;      // (nq) s_q
    const float* qamin;       // (nq) lower bound of sum_m min_c A[m][c]
    unsigned long long* counters;   // [0] survivor entries, [1] exact evaluations, [2] work items (may be null)
    const int* guard;               // fallback launches: run only if *guard != 0 (nullptr: always run)
    const int* qflag;               // fallback launches with (*guard & 3) == 0: only the queries with qflag[q] != 0
};

__host__ __device__ inline int ql_topk_cap(int k, int threads) {
    // TopK::flush's register reduction needs R * threads queue slots (R = 1, 2, 4 for k <= 32, 64, 128)
    const int r = k <= 32 ? 1 : k <= 64 ? 2 : 4;
    const int cap = r * threads;
    return cap < 512 ? 512 : cap;
}

inline bool ql_supported(int M, int d, int k) {
    (void)d;
    return (M == 16 || M == 32 || M == 64) && k <= 512;
}

// shared memory: [ table planes | residuals 4 x (dpad + 4) f32 | 4 x TopK | survivors | control | 2 work items ]
struct QlCtrl {            // 128 bytes
    int work;
    int nsurv[2];          // alternate from one drain to the next (zeroed a whole drain before reuse)
    int cold;
    float dis0[kQlQ];      // ||r_q||^2, accumulated by the warps that compute the residuals
    int qidx[kQlQ];        // the queries of the work item
    uint32_t pad_[20];
};
static_assert(sizeof(QlCtrl) == 128, "QlCtrl is 128 bytes");

__host__ __device__ inline int ql_res_stride(int d) { return ((d + 3) & ~3) + 4; }

template <int M>
__host__ __device__ inline size_t ql_smem_bytes(int d, int k) {
    return static_cast<size_t>(QlCfg<M>::kLutBytes) + sizeof(float) * kQlQ * ql_res_stride(d) +
           kQlQ * TopK::smem_bytes(k, ql_topk_cap(k, QlCfg<M>::kT)) + sizeof(uint32_t) * kQlSurvCap + sizeof(QlCtrl) +
           2 * sizeof(QlGroup);
}

// ------------------------------------------------------------------------------------------------------------------
// Index side: SB[j] = ||p_j||^2 + 2 (c_l - mu).p_j in float64, stored as 16 bits on a per-list grid, rounded down.
// One CTA per list (grid-stride); G[m][c] = ||pq[m][c]||^2 + 2 (c_l - mu)_m . pq[m][c] is tabulated in shared memory.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double ql_sb_of(const double* __restrict__ G, const uint8_t* __restrict__ code, int M) {
    double a = 0.0;
    if ((M & 15) == 0 && (reinterpret_cast<uintptr_t>(code) & 15) == 0) {
        for (int m0 = 0; m0 < M; m0 += 16) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(code + m0));
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 16; i++) a += G[(m0 + i) * 256 + ((w[i >> 2] >> (8 * (i & 3))) & 255u)];
        }
    } else {
        for (int m = 0; m < M; m++) a += G[m * 256 + __ldg(code + m)];
    }
    return a;
}

__global__ void __launch_bounds__(256) ql_sb_build_kernel(const float* __restrict__ cent, const float* __restrict__ pq,
                                                          const float* __restrict__ mu,
                                                          const int64_t* __restrict__ offsets,
                                                          const uint8_t* __restrict__ codes, int64_t nlist, int d, int M,
                                                          int dsub, uint16_t* __restrict__ snorm,
                                                          float* __restrict__ sbmin, float* __restrict__ sbstep) {
    extern __shared__ double ql_G[];   // [M][256]
    __shared__ double red_lo[8], red_hi[8];
    __shared__ float s_min, s_step;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int64_t list = blockIdx.x; list < nlist; list += gridDim.x) {
        const int64_t beg = offsets[list], n = offsets[list + 1] - beg;
        if (n <= 0) {   // uniform
            if (tid == 0) {
                sbmin[list] = 0.0f;
                sbstep[list] = 0.0f;
            }
            continue;
        }
        for (int e = tid; e < M * 256; e += 256) {
            const int m = e >> 8;
            const float* p = pq + static_cast<int64_t>(e) * dsub;
            double a = 0.0;
            for (int j = 0; j < dsub; j++) {
                const double pj = static_cast<double>(p[j]);
                const double cj = static_cast<double>(cent[list * d + m * dsub + j]) - static_cast<double>(mu[m * dsub + j]);
                a += pj * (pj + 2.0 * cj);
            }
            ql_G[e] = a;
        }
        __syncthreads();
        double lo = 1.0e300, hi = -1.0e300;
        for (int64_t i = tid; i < n; i += 256) {
            const double sb = ql_sb_of(ql_G, codes + (beg + i) * M, M);
            lo = fmin(lo, sb);
            hi = fmax(hi, sb);
        }
        for (int o = 16; o > 0; o >>= 1) {
            lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (lane == 0) {
            red_lo[wid] = lo;
            red_hi[wid] = hi;
        }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < 8; w++) {
                lo = fmin(lo, red_lo[w]);
                hi = fmax(hi, red_hi[w]);
            }
            float f = static_cast<float>(lo);
            if (static_cast<double>(f) > lo) f = nextafterf(f, -INFINITY);
            const double span = hi - static_cast<double>(f);
            float st = static_cast<float>(span / 65535.0);
            while (static_cast<double>(st) * 65535.0 < span) st = nextafterf(st, INFINITY);
            s_min = f;
            s_step = st;
            sbmin[list] = f;
            sbstep[list] = st;
        }
        __syncthreads();
        const double fmin64 = static_cast<double>(s_min), st64 = static_cast<double>(s_step);
        for (int64_t i = tid; i < n; i += 256) {
            const double sb = ql_sb_of(ql_G, codes + (beg + i) * M, M);
            long long u = st64 > 0.0 ? static_cast<long long>(floor((sb - fmin64) / st64)) : 0ll;
            u = u < 0 ? 0 : u > 65535 ? 65535 : u;
            while (u > 0 && fmin64 + st64 * static_cast<double>(u) > sb) u--;
            snorm[beg + i] = static_cast<uint16_t>(u);
        }
        __syncthreads();   // ql_G is rebuilt for the next list
    }
}

// mu = mean of the coarse centroids (float64 accumulation; any fixed vector is valid, this one keeps A and SB small)
__global__ void __launch_bounds__(256) ql_mean_kernel(const float* __restrict__ cent, int64_t nlist, int d,
                                                      float* __restrict__ mu) {
    __shared__ double red[256];
    const int j = blockIdx.x, tid = threadIdx.x;
    double a = 0.0;
    for (int64_t l = tid; l < nlist; l += 256) a += static_cast<double>(cent[l * d + j]);
    red[tid] = a;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) red[tid] += red[tid + o];
        __syncthreads();
    }
    if (tid == 0) mu[j] = static_cast<float>(red[0] / static_cast<double>(nlist));
}

// ------------------------------------------------------------------------------------------------------------------
// Query side: A_q[m][c] = -2 (q - mu)_m . pq[m][c] for QB = 64 / M queries per CTA (thread = code value c, the
// codebook row is loaded once for all of them), quantised with the query's own scale.  Offsets and scale come from
// the Cauchy-Schwarz bound |A[m][c]| <= B_m = 2 ||(q - mu)_m|| max_c ||pq[m][c]|| -- no reduction over the 256 code
// values is needed, at the price of at most one bit of the entries' range:
//   lut[q][c][m] = min(qmax, floor((A + B_m) * s_q (1 - 2^-20)))               s_q = qmax / (1.0001 max_m 2 B_m)
//   amin[q]      = - sum_m B_m  -  (dsub + 3) 2^-23 sum_m ||(q - mu)_m|| max_c ||pq[m][c]||  -  fl slack
// ------------------------------------------------------------------------------------------------------------------
template <int M>
__global__ void __launch_bounds__(256) ql_query_tables_kernel(const float* __restrict__ xq, int64_t nq,
                                                              const float* __restrict__ pq, const float* __restrict__ mu,
                                                              const float* __restrict__ pq_maxnorm, int d, int dsub,
                                                              uint32_t qmax, uint16_t* __restrict__ lut,
                                                              float* __restrict__ scale, float* __restrict__ amin) {
    constexpr int QB = 64 / M;
    const float kQ = static_cast<float>(qmax);
    extern __shared__ __align__(16) float ql_qc[];   // [QB][d]  q - mu
    __shared__ float s_B[QB][M], s_scale[QB];
    const int tid = threadIdx.x;
    const int64_t q0 = static_cast<int64_t>(blockIdx.x) * QB;
    for (int e = tid; e < QB * d; e += 256) {
        const int qb = e / d, j = e % d;
        const int64_t q = q0 + qb < nq ? q0 + qb : nq - 1;
        ql_qc[e] = __fsub_rn(xq[q * d + j], mu[j]);
    }
    __syncthreads();
    if (tid < QB * M) {
        const int qb = tid / M, m = tid % M;
        float nn = 0.0f;
        for (int j = 0; j < dsub; j++) nn = fmaf(ql_qc[qb * d + m * dsub + j], ql_qc[qb * d + m * dsub + j], nn);
        s_B[qb][m] = 2.0f * sqrtf(nn) * 1.0001f * pq_maxnorm[m];      // >= |A[m][c]| for every c
    }
    __syncthreads();
    if (tid < QB) {
        const int qb = tid;
        float range = 0.0f, sumB = 0.0f;
        for (int m = 0; m < M; m++) {
            range = fmaxf(range, 2.0f * s_B[qb][m]);
            sumB += s_B[qb][m];
        }
        // finite check: NaN / inf queries get scale 0 (nothing is filtered, the exact path decides)
        const bool ok = range > 0.0f && range < 1.0e30f && sumB < 1.0e30f;
        const float s = ok ? (kQ / (range * 1.0001f)) : 0.0f;
        s_scale[qb] = s;
        if (q0 + qb < nq) {
            scale[q0 + qb] = s;
            // rounding of the fp32 dot products ((dsub + 3) 2^-23 of |q_m| |p|, i.e. of B_m / 2) + of this sum
            const float slack = static_cast<float>(dsub + 3) * 1.1920929e-7f * sumB + sumB * 1.2e-7f * static_cast<float>(M);
            amin[q0 + qb] = ok ? -(sumB + slack) : 0.0f;
        }
    }
    __syncthreads();
    uint32_t w[QB][M / 2];
#pragma unroll
    for (int m = 0; m < M; m += 2) {
        float acc[QB][2];
#pragma unroll
        for (int qb = 0; qb < QB; qb++) acc[qb][0] = acc[qb][1] = 0.0f;
        const float* p0 = pq + (static_cast<int64_t>(m) * 256 + tid) * dsub;
        const float* p1 = p0 + 256 * dsub;
        if ((dsub & 3) == 0) {
            for (int j = 0; j < dsub; j += 4) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(p0 + j));
                const float4 b = __ldg(reinterpret_cast<const float4*>(p1 + j));
#pragma unroll
                for (int qb = 0; qb < QB; qb++) {
                    const float4 x = *reinterpret_cast<const float4*>(ql_qc + qb * d + m * dsub + j);
                    const float4 y = *reinterpret_cast<const float4*>(ql_qc + qb * d + (m + 1) * dsub + j);
                    acc[qb][0] = fmaf(x.w, a.w, fmaf(x.z, a.z, fmaf(x.y, a.y, fmaf(x.x, a.x, acc[qb][0]))));
                    acc[qb][1] = fmaf(y.w, b.w, fmaf(y.z, b.z, fmaf(y.y, b.y, fmaf(y.x, b.x, acc[qb][1]))));
                }
            }
        } else {
            for (int j = 0; j < dsub; j++) {
                const float a = __ldg(p0 + j), b = __ldg(p1 + j);
#pragma unroll
                for (int qb = 0; qb < QB; qb++) {
                    acc[qb][0] = fmaf(ql_qc[qb * d + m * dsub + j], a, acc[qb][0]);
                    acc[qb][1] = fmaf(ql_qc[qb * d + (m + 1) * dsub + j], b, acc[qb][1]);
                }
            }
        }
#pragma unroll
        for (int qb = 0; qb < QB; qb++) {
            const float s2 = s_scale[qb] * 0.999999f;
            const float x0 = (-2.0f * acc[qb][0] + s_B[qb][m]) * s2, x1 = (-2.0f * acc[qb][1] + s_B[qb][m + 1]) * s2;
            const uint32_t u0 = x0 > 0.0f ? min(static_cast<uint32_t>(__float2uint_rz(x0)), qmax) : 0u;
            const uint32_t u1 = x1 > 0.0f ? min(static_cast<uint32_t>(__float2uint_rz(x1)), qmax) : 0u;
            w[qb][m / 2] = u0 | (u1 << 16);
        }
    }
#pragma unroll
    for (int qb = 0; qb < QB; qb++) {
        if (q0 + qb >= nq) break;
        uint4* dst = reinterpret_cast<uint4*>(lut + ((q0 + qb) * 256 + tid) * M);
#pragma unroll
        for (int i = 0; i < M / 8; i++) dst[i] = make_uint4(w[qb][4 * i], w[qb][4 * i + 1], w[qb][4 * i + 2], w[qb][4 * i + 3]);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// pair setup: the counting sort of kernels.cuh with the key (rank bucket, list), emitting QlGroup work items.
// Work items are handed out in key order, so every query's NEAREST list (rank 0) is scanned first, ranks 1-3 next,
// the rest last: by the time the bulk of the pairs is scanned, every query has a k-th best distance that is close to
// the final one, and the filter removes 3-4 times more codes than with the pairs in plain list order.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kQlBuckets = 3;
__device__ __forceinline__ int ql_bucket(int rank) { return rank == 0 ? 0 : rank < 4 ? 1 : 2; }

__global__ void ql_pair_hist_kernel(const int32_t* __restrict__ probe, int64_t npairs, int nprobe, int64_t nlist,
                                    int nbuckets, const int64_t* __restrict__ offsets, int* __restrict__ hist,
                                    PairStats* __restrict__ stats) {
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    unsigned long long codes = 0;
    if (i < npairs) {
        int l = probe[i];
        if (l >= 0) {
            int64_t sz = offsets[l + 1] - offsets[l];
            if (sz > 0) {
                atomicAdd(&hist[(nbuckets == 1 ? 0 : ql_bucket(static_cast<int>(i % nprobe))) * nlist + l], 1);
                codes = static_cast<unsigned long long>(sz);
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) codes += __shfl_down_sync(0xffffffffu, codes, o);
    if ((threadIdx.x & 31) == 0 && codes) atomicAdd(&stats->scan_codes, codes);
}

__global__ void ql_pair_scatter_kernel(const int32_t* __restrict__ probe, int64_t npairs, int nprobe, int64_t nlist,
                                       int nbuckets, const int64_t* __restrict__ offsets, const int* __restrict__ start,
                                       const int* __restrict__ gstart, int* __restrict__ cursor,
                                       int32_t* __restrict__ order, QlGroup* __restrict__ groups, int gsz) {
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= npairs) return;
    int l = probe[i];
    if (l < 0) return;
    const int64_t beg = offsets[l], sz = offsets[l + 1] - beg;
    if (sz <= 0) return;
    const int64_t key = (nbuckets == 1 ? 0 : ql_bucket(static_cast<int>(i % nprobe))) * nlist + l;
    const int rank = atomicAdd(&cursor[key], 1);
    order[start[key] + rank] = static_cast<int32_t>(i);
    // gsz = 4: four pairs per work item; gsz = 2: two (slots 2, 3 stay -1: the two-query filter, scan_stream.cuh)
    const int slot = gsz == 4 ? (rank & 3) : (rank & 1);
    QlGroup* g = groups + gstart[key] + (gsz == 4 ? (rank >> 2) : (rank >> 1));
    g->pair[slot] = static_cast<int32_t>(i);
    g->query[slot] = static_cast<int32_t>(i / nprobe);
    if (slot == 0) {
        g->list = l;
        g->n = static_cast<uint32_t>(sz);
        g->beg = beg;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// the scan
// ------------------------------------------------------------------------------------------------------------------
template <int M>
struct QlCode {
    uint4 v[M / 16];
    uint32_t s;   // the vector's 16-bit SB grid index
};

template <int M>
__device__ __forceinline__ QlCode<M> ql_load_code(const uint4* __restrict__ lp, const uint16_t* __restrict__ sp,
                                                  uint32_t idx, uint32_t n) {
    QlCode<M> c;
#pragma unroll
    for (int h = 0; h < M / 16; h++) c.v[h] = make_uint4(0u, 0u, 0u, 0u);
    c.s = 0u;
    if (idx < n) {
#pragma unroll
        for (int h = 0; h < M / 16; h++) c.v[h] = __ldg(lp + static_cast<size_t>(idx) * (M / 16) + h);
        c.s = __ldg(sp + idx);
    }
    return c;
}

// Offsets of the 16 look-up steps, three per register with a zero top byte (the PRMT that builds the address takes
// its low byte from here and its two top bytes from the zero byte):  step p reads slot (r ^ p), r = lane % 16.
struct QlOffsets {
    uint32_t o[6];
};
__device__ __forceinline__ QlOffsets ql_make_offsets(int r) {
    QlOffsets f;
#pragma unroll
    for (int i = 0; i < 6; i++) {
        uint32_t v = 0u;
#pragma unroll
        for (int b = 0; b < 3; b++) {
            const int p = 3 * i + b;
            if (p < 16) v |= static_cast<uint32_t>((r ^ p) * 8) << (8 * b);
        }
        f.o[i] = v;
    }
    return f;
}

// (c << 8) | 8 * (r ^ P): byte 0 = offset byte (P % 3) of o[P / 3], byte 1 = byte B of w, bytes 2, 3 = the zero byte
template <int B, int P>
__device__ __forceinline__ uint2 ql_lookup(const char* __restrict__ lutb, uint32_t w, const QlOffsets& f) {
    const uint32_t a = __byte_perm(w, f.o[P / 3], 0x7700u | (B << 4) | (4 + P % 3));
    return *reinterpret_cast<const uint2*>(lutb + a);
}

// The 16 bytes of one chunk, permuted so that byte p of the result is byte (p ^ r) of the code: the lane walks its own
// code in the order m = r ^ 0, r ^ 1, ...; integer sums do not care, and the 16 lanes of an LDS.64 phase read 16
// distinct slots of (possibly different) rows -- all 32 banks, no conflict, for ANY code bytes.
__device__ __forceinline__ uint2 ql_block16(const char* __restrict__ lutb, const uint4& code, bool x8, bool x4,
                                            uint32_t bsel, const QlOffsets& f) {
    const uint32_t y0 = x8 ? code.z : code.x, y1 = x8 ? code.w : code.y, y2 = x8 ? code.x : code.z,
                   y3 = x8 ? code.y : code.w;
    const uint32_t z0 = x4 ? y1 : y0, z1 = x4 ? y0 : y1, z2 = x4 ? y3 : y2, z3 = x4 ? y2 : y3;
    const uint32_t w0 = __byte_perm(z0, 0u, bsel), w1 = __byte_perm(z1, 0u, bsel), w2 = __byte_perm(z2, 0u, bsel),
                   w3 = __byte_perm(z3, 0u, bsel);
    uint32_t s01 = 0u, s23 = 0u;
#define QL_STEP(W, B, P)                                   \
    {                                                      \
        const uint2 t = ql_lookup<B, P>(lutb, W, f);       \
        s01 += t.x;                                        \
        s23 += t.y;                                        \
    }
    QL_STEP(w0, 0, 0) QL_STEP(w0, 1, 1) QL_STEP(w0, 2, 2) QL_STEP(w0, 3, 3)
    QL_STEP(w1, 0, 4) QL_STEP(w1, 1, 5) QL_STEP(w1, 2, 6) QL_STEP(w1, 3, 7)
    QL_STEP(w2, 0, 8) QL_STEP(w2, 1, 9) QL_STEP(w2, 2, 10) QL_STEP(w2, 3, 11)
    QL_STEP(w3, 0, 12) QL_STEP(w3, 1, 13) QL_STEP(w3, 2, 14) QL_STEP(w3, 3, 15)
#undef QL_STEP
    return make_uint2(s01, s23);
}

// Two queries per work item (lists probed by one or two queries: small batches x nprobe): a table entry is ONE 32-bit
// word (u_q0 | u_q1 << 16), so a look-up is an LDS.32 -- one shared-memory wavefront per warp instead of the two an LDS.64
// takes, which the four-query layout wastes on empty query slots there.  Row = code value (256-byte stride as above),
// word (lane / 16) * 16 + (r ^ step): the table is stored twice per row so that the two half-warps -- which walk
// different codes -- use disjoint banks: 32 distinct banks for ANY code bytes.
__device__ __forceinline__ QlOffsets ql_make_offsets_two(int lane) {
    const int r = lane & 15, half = lane >> 4;
    QlOffsets f;
#pragma unroll
    for (int i = 0; i < 6; i++) {
        uint32_t v = 0u;
#pragma unroll
        for (int b = 0; b < 3; b++) {
            const int p = 3 * i + b;
            if (p < 16) v |= static_cast<uint32_t>(half * 64 + (r ^ p) * 4) << (8 * b);
        }
        f.o[i] = v;
    }
    return f;
}
template <int B, int P>
__device__ __forceinline__ uint32_t ql_lookup_two(const char* __restrict__ lutb, uint32_t w, const QlOffsets& f) {
    const uint32_t a = __byte_perm(w, f.o[P / 3], 0x7700u | (B << 4) | (4 + P % 3));
    return *reinterpret_cast<const uint32_t*>(lutb + a);
}
// packed lower bounds (q0 | q1 << 16) of one 16-byte code; each half <= 16 * 2047 < 2^16: no carry between the halves
__device__ __forceinline__ uint32_t ql_block16_two(const char* __restrict__ lutb, const uint4& code, bool x8, bool x4,
                                                   uint32_t bsel, const QlOffsets& f) {
    const uint32_t y0 = x8 ? code.z : code.x, y1 = x8 ? code.w : code.y, y2 = x8 ? code.x : code.z,
                   y3 = x8 ? code.y : code.w;
    const uint32_t z0 = x4 ? y1 : y0, z1 = x4 ? y0 : y1, z2 = x4 ? y3 : y2, z3 = x4 ? y2 : y3;
    const uint32_t w0 = __byte_perm(z0, 0u, bsel), w1 = __byte_perm(z1, 0u, bsel), w2 = __byte_perm(z2, 0u, bsel),
                   w3 = __byte_perm(z3, 0u, bsel);
    uint32_t sa = 0u, sb = 0u;   // two chains
#define QL_STEP2(S, W, B, P) S += ql_lookup_two<B, P>(lutb, W, f);
    QL_STEP2(sa, w0, 0, 0) QL_STEP2(sb, w0, 1, 1) QL_STEP2(sa, w0, 2, 2) QL_STEP2(sb, w0, 3, 3)
    QL_STEP2(sa, w1, 0, 4) QL_STEP2(sb, w1, 1, 5) QL_STEP2(sa, w1, 2, 6) QL_STEP2(sb, w1, 3, 7)
    QL_STEP2(sa, w2, 0, 8) QL_STEP2(sb, w2, 1, 9) QL_STEP2(sa, w2, 2, 10) QL_STEP2(sb, w2, 3, 11)
    QL_STEP2(sa, w3, 0, 12) QL_STEP2(sb, w3, 1, 13) QL_STEP2(sa, w3, 2, 14) QL_STEP2(sb, w3, 3, 15)
#undef QL_STEP2
    return sa + sb;
}

__device__ __forceinline__ void ql_copy_group_async(QlGroup* dst, const QlGroup* src) {
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
    const char* s = reinterpret_cast<const char*>(src);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(s) : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d + 16), "l"(s + 16) : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d + 32), "l"(s + 32) : "memory");
}

// The four packed lower bounds of all chunks of a code -> four "2^23 + integer" floats.  One chunk: the packed
// half-words go straight into the float's mantissa (one PRMT each).  Several chunks (M = 32, 64): every chunk's packed
// sums (< 2^15 each) are widened and added in 32 bits (M * 2047 < 2^23).
template <int M>
__device__ __forceinline__ void ql_bounds_as_floats(const char* __restrict__ lutb, const uint4 (&v)[M / 16], bool x8, bool x4,
                                                    uint32_t bsel, const QlOffsets& offs, float (&f)[4]) {
    if constexpr (M == 16) {
        const uint2 lb = ql_block16(lutb, v[0], x8, x4, bsel, offs);
        f[0] = __uint_as_float(__byte_perm(lb.x, 0x4b000000u, 0x7410));
        f[1] = __uint_as_float(__byte_perm(lb.x, 0x4b000000u, 0x7432));
        f[2] = __uint_as_float(__byte_perm(lb.y, 0x4b000000u, 0x7410));
        f[3] = __uint_as_float(__byte_perm(lb.y, 0x4b000000u, 0x7432));
    } else {
        uint32_t s0 = 0u, s1 = 0u, s2 = 0u, s3 = 0u;
#pragma unroll
        for (int h = 0; h < M / 16; h++) {
            const uint2 lb = ql_block16(lutb + (h >> 1) * kQlPlaneBytes + (h & 1) * 128, v[h], x8, x4, bsel, offs);
            s0 += lb.x & 0xffffu;
            s1 += lb.x >> 16;
            s2 += lb.y & 0xffffu;
            s3 += lb.y >> 16;
        }
        f[0] = __uint_as_float(0x4b000000u | s0);
        f[1] = __uint_as_float(0x4b000000u | s1);
        f[2] = __uint_as_float(0x4b000000u | s2);
        f[3] = __uint_as_float(0x4b000000u | s3);
    }
}

// threshold constant of one query in the magic-number domain: a code survives iff !(2^23 + LB > fma(v, -astep, b))
__device__ __forceinline__ float ql_threshold_const(uint32_t thr_bits, float scale, float base, float mag) {
    const float thr = __uint_as_float(thr_bits);
    const float y = (thr + base) + 4.8e-7f * (fabsf(thr) + mag);
    const float sy = scale * y;
    return (sy + fabsf(sy) * 4.8e-7f) + (2.0f + kQlMagic);   // inf / NaN stay inf / NaN: everything survives
}

// exact distance of one code for one query, the oracle's operation order (j ascending inside a sub-quantizer, then m
// ascending), every operation rounded separately.  rq: the query's residual in shared memory; vector loads where the
// sub-vector length allows.
template <int M>
__device__ __forceinline__ float ql_exact(const uint4 (&cc)[M / 16], const float* __restrict__ rq,
                                          const float* __restrict__ pq, int dsub) {
    float acc = 0.0f;
#pragma unroll
    for (int h = 0; h < M / 16; h++) {
#pragma unroll 1
        for (int w = 0; w < 4; w++) {
            const uint32_t cw = w == 0 ? cc[h].x : w == 1 ? cc[h].y : w == 2 ? cc[h].z : cc[h].w;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int m = 16 * h + 4 * w + b;
                const uint32_t c = (cw >> (8 * b)) & 255u;
                const float* pc = pq + (static_cast<int64_t>(m) * 256 + c) * dsub;
                const float* rr = rq + m * dsub;
                float t = 0.0f;
                if ((dsub & 3) == 0) {
                    for (int j = 0; j < dsub; j += 4) {
                        const float4 pv = __ldg(reinterpret_cast<const float4*>(pc + j));
                        const float4 rv = *reinterpret_cast<const float4*>(rr + j);
                        t = sqdiff_acc(t, rv.x, pv.x);
                        t = sqdiff_acc(t, rv.y, pv.y);
                        t = sqdiff_acc(t, rv.z, pv.z);
                        t = sqdiff_acc(t, rv.w, pv.w);
                    }
                } else if ((dsub & 1) == 0) {
                    for (int j = 0; j < dsub; j += 2) {
                        const float2 pv = __ldg(reinterpret_cast<const float2*>(pc + j));
                        const float2 rv = *reinterpret_cast<const float2*>(rr + j);
                        t = sqdiff_acc(t, rv.x, pv.x);
                        t = sqdiff_acc(t, rv.y, pv.y);
                    }
                } else {
                    for (int j = 0; j < dsub; j++) t = sqdiff_acc(t, rr[j], __ldg(pc + j));
                }
                acc = __fadd_rn(acc, t);
            }
        }
    }
    return acc;
}

// Fold the pending candidates of all Q queues into their best lists.  One out-of-line copy per kernel: the sorting
// networks are large and run once or twice per work item.  All kT threads call, after a barrier.
template <int THREADS>
__device__ __noinline__ void ql_fold_all(unsigned char* tk_base, size_t tk_bytes, int k, int cap, uint4 ext4) {
    TopK tk[kQlQ];
    const uint32_t ext[kQlQ] = {ext4.x, ext4.y, ext4.z, ext4.w};
#pragma unroll
    for (int q = 0; q < kQlQ; q++) tk[q].bind(tk_base + q * tk_bytes, k, cap);
    if (!topk_fold_small<THREADS, kQlQ>(tk, ext)) {
#pragma unroll 1
        for (int q = 0; q < kQlQ; q++) tk[q].template flush<THREADS>(ext[q]);
    }
}

template <int M>
__global__ void __launch_bounds__(QlCfg<M>::kT, M == 64 ? 1 : 2)   // M = 64: one CTA per SM (128 KB of tables), 255 registers
scan_qlut_kernel(const ScanParams p, const QlParams ql) {
    using Cfg = QlCfg<M>;
    constexpr int kT = Cfg::kT;
    constexpr int Q = kQlQ;
    extern __shared__ __align__(1024) unsigned char smem_ql[];
    char* lutb = reinterpret_cast<char*>(smem_ql);
    const int rstride = ql_res_stride(p.d);
    float* res = reinterpret_cast<float*>(smem_ql + Cfg::kLutBytes);                      // [Q][rstride]
    unsigned char* tk_base = reinterpret_cast<unsigned char*>(res + Q * rstride);
    const int kCap = ql_topk_cap(p.k, kT);
    const size_t tk_bytes = TopK::smem_bytes(p.k, kCap);
    TopK tk[Q];
#pragma unroll
    for (int q = 0; q < Q; q++) tk[q].bind(tk_base + q * tk_bytes, p.k, kCap);
    uint32_t* surv = reinterpret_cast<uint32_t*>(tk_base + Q * tk_bytes);
    QlCtrl* ctrl = reinterpret_cast<QlCtrl*>(surv + kQlSurvCap);
    QlGroup* s_grp = reinterpret_cast<QlGroup*>(ctrl + 1);

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int r = lane & 15;
    const QlOffsets offs = ql_make_offsets(r);
    const bool x8 = (r & 8) != 0, x4 = (r & 4) != 0;
    // byte (p ^ (r & 3)) of a word: selectors 3210, 2301, 1032, 0123
    const uint32_t bsel = (r & 3) == 0 ? 0x3210u : (r & 3) == 1 ? 0x2301u : (r & 3) == 2 ? 0x1032u : 0x0123u;
    const int ovf = ql.guard ? *reinterpret_cast<const volatile int*>(ql.guard) : 3;
    if (ovf == 0) return;                                                            // fallback launch, not needed
    const int ngroups = p.stats->ngroups;
    const int dsub = p.dsub;
    int* const work_counter = ql.guard ? &p.stats->work_counter2 : &p.stats->work_counter;
    // cold start: that many survivors are evaluated before the first thresholds exist
    const int kBoot = min(kT, (p.k + 31) & ~31);
    unsigned long long n_surv = 0ull, n_exact = 0ull, n_items = 0ull;   // thread 0's copies are reported

    int next_work = 0, buf = 0;
    if (tid == 0) {
        ctrl->nsurv[0] = 0;
        ctrl->nsurv[1] = 0;
        next_work = atomicAdd(work_counter, 1);
        if (next_work < ngroups) ql_copy_group_async(&s_grp[0], static_cast<const QlGroup*>(p.groups) + next_work);
    }
    for (;;) {
        if (tid == 0) {
            ctrl->work = next_work;
#pragma unroll
            for (int q = 0; q < Q; q++) ctrl->dis0[q] = 0.0f;
            asm volatile("cp.async.wait_all;" ::: "memory");
        }
        __syncthreads();
        const int wk = ctrl->work;
        if (wk >= ngroups) break;
        if (tid == 0) next_work = atomicAdd(work_counter, 1);
        const QlGroup grp = s_grp[buf];
        int pair[Q], qi[Q];
        uint32_t vmask = 0u;
#pragma unroll
        for (int q = 0; q < Q; q++) {
            const bool has = grp.pair[q] >= 0;
            pair[q] = has ? grp.pair[q] : grp.pair[0];
            qi[q] = has ? grp.query[q] : grp.query[0];
            vmask |= has ? (1u << q) : 0u;
        }
        if ((ovf & 3) == 0) {
            // fallback after slab overflows: only the flagged queries are recomputed (CTA-uniform decision)
#pragma unroll
            for (int q = 0; q < Q; q++)
                if (((vmask >> q) & 1u) && __ldg(ql.qflag + qi[q]) == 0) vmask &= ~(1u << q);
            if (vmask == 0u) {
                buf ^= 1;
                if (tid == 0 && next_work < ngroups)
                    ql_copy_group_async(&s_grp[buf], static_cast<const QlGroup*>(p.groups) + next_work);
                continue;
            }
        }
        const int list = grp.list;
        const uint32_t n = grp.n;
        const uint4* lp = reinterpret_cast<const uint4*>(p.codes + grp.beg * M);
        const uint16_t* sp = ql.snorm + grp.beg;

        // ONE thread reads each query's current threshold (other CTAs lower it concurrently): every thread must see
        // the same value, the control flow below depends on it
#pragma unroll
        for (int q = 0; q < Q; q++)
            if (tid == q) {
                tk[q].reset(*reinterpret_cast<volatile uint32_t*>(p.qthr + qi[q]));
                ctrl->qidx[q] = qi[q];
            }

        // a2: residuals r = fl(q - c), kept for the exact evaluation of the survivors; ||r||^2 summed on the way (any
        // order: only a bound is needed)
        {
            float n2[Q];
#pragma unroll
            for (int q = 0; q < Q; q++) n2[q] = 0.0f;
            for (int j = tid; j < p.d; j += kT) {
                const float cj = __ldg(p.cent + static_cast<int64_t>(list) * p.d + j);
#pragma unroll
                for (int q = 0; q < Q; q++) {
                    const float rj = __fsub_rn(__ldg(p.xq + static_cast<int64_t>(qi[q]) * p.d + j), cj);
                    res[q * rstride + j] = rj;
                    n2[q] = fmaf(rj, rj, n2[q]);
                }
            }
            if (wid * 32 < p.d) {   // warps that hold a part of the residuals
#pragma unroll
                for (int q = 0; q < Q; q++) {
                    for (int o = 16; o > 0; o >>= 1) n2[q] += __shfl_xor_sync(0xffffffffu, n2[q], o);
                    if (lane == 0) atomicAdd(&ctrl->dis0[q], n2[q]);
                }
            }
        }
        // the four queries' tables, interleaved entry-wise: row c of chunk h = 16 slots x (u_q0, u_q1, u_q2, u_q3).
        // Eight lanes cover the eight slot pairs of one row (128 contiguous bytes: conflict-free STS.128), four rows
        // per warp and step; eight steps' loads are in flight together.
        {
            const int sp2 = lane & 7;   // slots 2 sp2, 2 sp2 + 1
            constexpr int kSteps = Cfg::kChunks * 64 / (kT / 32);   // per warp
#pragma unroll 1
            for (int i0 = 0; i0 < kSteps; i0 += 8) {
                uint32_t a[8][Q];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int it = wid + (i0 + i) * (kT / 32);
                    const int h = it >> 6, row = ((it & 63) << 2) + (lane >> 3);
#pragma unroll
                    for (int q = 0; q < Q; q++)
                        a[i][q] = __ldg(reinterpret_cast<const uint32_t*>(ql.qlut + (static_cast<int64_t>(qi[q]) * 256 + row) * M) +
                                        h * 8 + sp2);
                }
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int it = wid + (i0 + i) * (kT / 32);
                    const int h = it >> 6, row = ((it & 63) << 2) + (lane >> 3);
                    uint4 e;
                    e.x = __byte_perm(a[i][0], a[i][1], 0x5410);
                    e.y = __byte_perm(a[i][2], a[i][3], 0x5410);
                    e.z = __byte_perm(a[i][0], a[i][1], 0x7632);
                    e.w = __byte_perm(a[i][2], a[i][3], 0x7632);
                    *reinterpret_cast<uint4*>(lutb + (h >> 1) * kQlPlaneBytes + row * 256 + (h & 1) * 128 + sp2 * 16) = e;
                }
            }
        }
        // pair constants (loads issued before the barrier)
        float c_am[Q], c_s[Q];
#pragma unroll
        for (int q = 0; q < Q; q++) {
            c_am[q] = __ldg(ql.qamin + qi[q]);
            c_s[q] = __ldg(ql.qscale + qi[q]);
        }
        const float c_sm = __ldg(ql.sbmin + list), c_st = __ldg(ql.sbstep + list);
        __syncthreads();
        float c_base[Q], c_mag[Q];
#pragma unroll
        for (int q = 0; q < Q; q++) {
            const float a = ctrl->dis0[q];
            const float dis0 = a * (1.0f - static_cast<float>(p.d + 5) * 5.9604645e-8f);
            const float rn = sqrtf(a) * 1.00001f + ql.pmax;
            const float E = static_cast<float>(dsub + M + 8) * 5.9604645e-8f * rn * rn * 1.00001f;
            c_mag[q] = fabsf(E) + fabsf(dis0) + fabsf(c_am[q]) + fabsf(c_sm);
            c_base[q] = (((E - dis0) - c_am[q]) - c_sm) + 4.8e-7f * c_mag[q];
        }
        buf ^= 1;
        if (tid == 0 && next_work < ngroups)
            ql_copy_group_async(&s_grp[buf], static_cast<const QlGroup*>(p.groups) + next_work);

        float na[Q], tb[Q];   // -astep and the threshold constant of every query (per-thread copies)
        uint32_t th[Q], ext[Q];
        bool cold = false;
#pragma unroll
        for (int q = 0; q < Q; q++) {
            na[q] = -(c_s[q] * c_st * 0.999999f);
            th[q] = tk[q].threshold();
            ext[q] = th[q];
            tb[q] = (vmask >> q) & 1u ? ql_threshold_const(th[q], c_s[q], c_base[q], c_mag[q]) : -INFINITY;
            cold = cold || (((vmask >> q) & 1u) && th[q] >= kInfBits);
        }
        int sphase = 0;

        // the filter on one code: packed lower bounds, then "2^23 + LB > threshold constant - astep * v" per query
        auto test = [&](const QlCode<M>& c) -> uint32_t {
            float f[4];
            ql_bounds_as_floats<M>(lutb, c.v, x8, x4, bsel, offs, f);
            const float vs = static_cast<float>(c.s);
            const bool h0 = !(f[0] > fmaf(vs, na[0], tb[0])), h1 = !(f[1] > fmaf(vs, na[1], tb[1])),
                       h2 = !(f[2] > fmaf(vs, na[2], tb[2])), h3 = !(f[3] > fmaf(vs, na[3], tb[3]));
            return (h0 ? 1u : 0u) | (h1 ? 2u : 0u) | (h2 ? 4u : 0u) | (h3 ? 8u : 0u);
        };
        auto enqueue = [&](bool hit, uint32_t entry) {
            const unsigned bal = __ballot_sync(0xffffffffu, hit);
            if (bal) {
                int slot = 0;
                const int leader = __ffs(bal) - 1;
                if (lane == leader) slot = atomicAdd(&ctrl->nsurv[sphase], __popc(bal));
                slot = __shfl_sync(0xffffffffu, slot, leader) + __popc(bal & lanemask_lt());
                if (hit) surv[slot] = entry;
            }
        };
        // after a fold (called by all threads, right after a barrier): publish this item's k-th best distances, pick up
        // what other CTAs found for the same queries, recompute the filter constants
        auto refresh = [&]() {
            if (tid < Q && ((vmask >> tid) & 1u)) {
                TopK t;
                t.bind(tk_base + tid * tk_bytes, p.k, kCap);
                const uint32_t mine = t.threshold();
                uint32_t* g = p.qthr + ctrl->qidx[tid];
                const uint32_t seen = t.count() == p.k ? atomicMin(g, mine) : *reinterpret_cast<volatile uint32_t*>(g);
                if (seen < mine) t.meta[3] = static_cast<int>(seen);
            }
            __syncthreads();
            bool c = false;
#pragma unroll
            for (int q = 0; q < Q; q++) {
                th[q] = tk[q].threshold();
                ext[q] = th[q];
                tb[q] = (vmask >> q) & 1u ? ql_threshold_const(th[q], c_s[q], c_base[q], c_mag[q]) : -INFINITY;
                c = c || (((vmask >> q) & 1u) && th[q] >= kInfBits);
            }
            cold = c;
        };
        auto fold_all = [&]() { ql_fold_all<kT>(tk_base, tk_bytes, p.k, kCap, make_uint4(ext[0], ext[1], ext[2], ext[3])); };
        // Exact evaluation of the queued survivors; folds the candidate queues and refreshes the thresholds.  Called by
        // all threads (CTA-uniform), right after a barrier.  While some query has no threshold yet, everything passes the
        // filter: then only kBoot entries are evaluated first, and the rest is filtered AGAIN with the thresholds that
        // gives (an exact evaluation costs as much as filtering several codes).
        auto drain = [&]() {
            const int ns = ctrl->nsurv[sphase];
            const bool boot = cold;
            if (tid == 0) n_surv += static_cast<unsigned long long>(ns);
            int base = 0, round = 0;
            while (base < ns) {
                const int width = (boot && round == 0) ? kBoot : kT;
                const int s = base + tid;
                if (tid < width && s < ns) {
                    const uint32_t e = surv[s];
                    const uint32_t idx = e >> 4;
                    uint32_t bits = e & 15u;
                    QlCode<M> c;
#pragma unroll
                    for (int h = 0; h < M / 16; h++) c.v[h] = __ldg(lp + static_cast<size_t>(idx) * (M / 16) + h);
                    if (boot && round > 0) {
                        c.s = __ldg(sp + idx);
                        bits &= test(c);
                    }
                    while (bits) {
                        const int q = __ffs(bits) - 1;
                        bits &= bits - 1u;
                        const uint32_t b = __float_as_uint(ql_exact<M>(c.v, res + q * rstride, p.pq, dsub));
                        unsigned char* tq = tk_base + q * tk_bytes;
                        int* meta = reinterpret_cast<int*>(tq + sizeof(uint64_t) * (2 * p.k + kCap));
                        if (b <= static_cast<uint32_t>(meta[3])) {
                            const int slot = atomicAdd(&meta[1], 1);
                            reinterpret_cast<uint64_t*>(tq)[2 * p.k + slot] = make_key(b, idx);
                        }
                        if (ql.counters) n_exact++;
                    }
                }
                base += width;
                round++;
                if (boot) {
                    __syncthreads();
                    fold_all();
                    refresh();
                } else {
                    // at most kT new entries per queue and round: fold when another round could overflow
                    bool over = false;
#pragma unroll
                    for (int q = 0; q < Q; q++) over = over || tk[q].pending() > kCap - kT;
                    if (__syncthreads_or(over) && base < ns) fold_all();
                }
            }
            if (tid == 0) ctrl->nsurv[sphase] = 0;   // next used after the NEXT drain: several barriers from now
            sphase ^= 1;
            if (!boot) {
                fold_all();
                refresh();
            }
        };

        // a4: the filter.  Block b = codes b*kT + tid.  Chunks of eight blocks (one while some query has no threshold
        // yet), then a CTA-wide check: drain once enough survivors are waiting (tight thresholds early are worth more
        // than fewer drains), and always after the last block.
        const uint32_t nblk = (n + kT - 1) / kT;
        uint32_t blk = 0;
        bool primed = false;
        QlCode<M> c0, c1, c2, c3;
#define QL_ITER(CUR, LOADTO, TB)                                                                     \
    {                                                                                                \
        LOADTO = ql_load_code<M>(lp, sp, base + (TB + 2) * kT, n);                                   \
        const uint32_t idx = base + TB * kT;                                                         \
        const uint32_t m_ = test(CUR);                                                               \
        enqueue(idx < n && m_ != 0u, (idx << 4) | m_);                                               \
    }
#pragma unroll 1
        while (blk < nblk) {
            if (cold) {
                const uint32_t idx = blk * kT + tid;
                const QlCode<M> c = ql_load_code<M>(lp, sp, idx, n);
                const uint32_t m = test(c);
                enqueue(idx < n && m != 0u, (idx << 4) | m);
                blk++;
                primed = false;
            } else {
                if (!primed) {
                    c0 = ql_load_code<M>(lp, sp, blk * kT + tid, n);
                    c1 = ql_load_code<M>(lp, sp, (blk + 1) * kT + tid, n);
                    primed = true;
                }
#pragma unroll 1
                for (int rnd = 0; rnd < 2 && blk < nblk; rnd++, blk += 4) {
                    const uint32_t base = blk * kT + tid;
                    QL_ITER(c0, c2, 0)
                    if (blk + 1 < nblk) QL_ITER(c1, c3, 1)
                    if (blk + 2 < nblk) QL_ITER(c2, c0, 2)
                    if (blk + 3 < nblk) QL_ITER(c3, c1, 3)
                }
            }
            const int seen = *reinterpret_cast<volatile int*>(&ctrl->nsurv[sphase]);
            if (__syncthreads_or(seen > (cold ? 0 : p.quad_drain_at)) || blk >= nblk) drain();
        }
#undef QL_ITER
#pragma unroll
        for (int q = 0; q < Q; q++) {
            if (vmask & (1u << q)) {
                const int nb = tk[q].count();
                const uint64_t* s = tk[q].sorted();
                for (int i = tid; i < nb; i += kT) p.out_keys[static_cast<int64_t>(pair[q]) * p.k + i] = s[i];
                if (tid == 0) p.out_cnt[pair[q]] = nb;
            }
        }
        if (tid == 0) n_items++;
        // no barrier here: the one at the top of the next item separates these reads from its first writes
    }
    if (ql.counters) {
        // per-thread exact-evaluation counts -> one atomic per warp
        for (int o = 16; o > 0; o >>= 1) n_exact += __shfl_xor_sync(0xffffffffu, n_exact, o);
        if (lane == 0 && n_exact) atomicAdd(ql.counters + 1, n_exact);
        if (tid == 0) {
            atomicAdd(ql.counters + 0, n_surv);
            atomicAdd(ql.counters + 2, n_items);
        }
    }
}

}  // namespace b200
