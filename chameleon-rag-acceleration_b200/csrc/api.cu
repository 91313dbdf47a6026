// api.cu -- C-ABI (include/b200_ivfpq.h) and host-side orchestration of the search path.
// One index handle owns only its workspace; codebooks and inverted lists are borrowed device pointers
// (torch tensors on the Python side), so a populated index costs no extra HBM.
#include <array>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/b200_ivfpq.h"
#include "kernels.cuh"
#include "scan_skew.cuh"
#include "scan_duo.cuh"
#include "scan_duo32.cuh"
#include "scan_quad.cuh"
#include "coarse_tc.cuh"
#include "select_radix.cuh"
#include "coarse_small.cuh"
#include "qlut_api.h"

using namespace b200;

namespace {

thread_local std::string g_last_error;
std::atomic<int64_t> g_launches{0};
std::atomic<uint64_t> g_ws_epoch{0};   // bumped whenever a workspace buffer is (re)allocated: invalidates CUDA graphs

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t err__ = (expr);                                                                 \
        if (err__ != cudaSuccess)                                                                   \
            return fail(B200_IVFPQ_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), \
                        __FILE__, __LINE__);                                                        \
    } while (0)

#define LAUNCH_CHECK()                 \
    do {                               \
        g_launches.fetch_add(1);       \
        CUDA_TRY(cudaGetLastError());  \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        g_ws_epoch.fetch_add(1);
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(B200_IVFPQ_ENOMEM, "workspace cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        }
        cap = want;
        return 0;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T* as() const {
        return reinterpret_cast<T*>(p);
    }
};

constexpr size_t kCoarseMatrixBudget = size_t(3) << 30;   // bytes of coarse distances per query chunk
constexpr size_t kPairOutBudget = size_t(2) << 30;        // bytes of per-pair candidates per query chunk
// The four-query filter kernel halves the per-code cost of the two-query kernel but has the higher cost per work
// item (four LUTs, exact LUT to global memory, survivor drains): it wins once lists are a few thousand codes long
// (measured on the C2 shape: 12.2k-code lists 9.5 vs 12.5 ms, 1.5k-code lists 4.9 vs 4.1 ms).
constexpr int64_t kQuadMinAvgList = 4096;

}  // namespace

struct b200_ivfpq_index {
    int d = 0, M = 0, nbits = 8, dsub = 0;
    int64_t nlist = 0, ntotal = 0;
    const float* cent = nullptr;
    const float* pq = nullptr;
    const uint8_t* codes = nullptr;
    const int64_t* ids = nullptr;
    bool has_lists = false;
    int device = 0, num_sms = 148;
    int scan_variant = 0;   // 0 = auto, 1 = generic, 2 = skewed (conflict-free), 3 = two-query skewed (scan_duo.cuh),
                            // 4 = four-query integer filter + exact survivors (scan_quad.cuh), 5 = per-query-table filter (scan_qlut.cuh)
    int64_t max_list = 0;   // longest inverted list
    int64_t nonempty = 0;   // lists holding at least one entry (a by-list shard leaves the others empty)
    int force_nseg = 0;     // B200_IVFPQ_NSEG: override the list segmentation (tests)
    int quad_drain_at = 256;   // B200_IVFPQ_QUAD_DRAIN: survivors queued per query slot before the exact evaluation runs
    // per-query-table filter scan (scan_qlut.cuh): per-index data (built lazily by ql_prepare) + per-batch tables
    DevBuf ql_mu, ql_snorm, ql_sbmin, ql_sbstep, ql_lut, ql_scale, ql_amin, ql_counters;
    // streaming pipeline (scan_stream.cuh)
    DevBuf st_srec, st_sfill, st_ctr, st_slab, st_qcnt, st_qflag, st_qkey, st_prefix, st_pdis;
    // a search split in two for the multi-GPU threshold exchange (b200_ivfpq_search_preassigned_begin / _finish)
    struct Pending {
        bool valid = false;
        ScanParams sp;
        QlHostParams qp;
        StHostBuffers sb;
        int64_t nq = 0;
        int k = 0, nprobe = 0, st_ctas = 0, ql_ctas = 0;
        cudaStream_t st = nullptr;
    } pend;
    // b200_ivfpq_prepare_queries: the per-query tables of the NEXT search of (xq, nq) are already being built on `side`
    struct Prepared {
        bool valid = false;
        const float* xq = nullptr;
        int64_t nq = 0;
    } prep;
    cudaStream_t side = nullptr;              // per-query tables overlap the coarse stage on this stream
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool side_ok() {
        if (side) return true;
        if (cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            side = nullptr;
            return false;
        }
        return true;
    }
    int st_mode = 1;            // B200_IVFPQ_STREAM=0: in-kernel top-k (scan_qlut_kernel) instead of the streaming pipeline
    double st_rate = 0.01;      // B200_IVFPQ_STREAM_RATE: survivor records provisioned per (query, code) pair
    int st_two = -1;            // B200_IVFPQ_STREAM_TWO: -1 auto (two-query LDS.32 filter when npairs <= nlist, M = 16),
                                // 0 never, 1 always, 2 = the two-query filter with bulk-async code tiles: measured SLOWER
                                // than the four-query kernel everywhere (profiles/r2_sweep_c2_two_query_filter.json)
    int st_capq = 0;            // B200_IVFPQ_STREAM_CAPQ: keys per query slab (0 = max(1024, 32 k))
    double st_minrec = 4.0 * 1048576.0;   // B200_IVFPQ_STREAM_MINREC: lower bound of the record buffer (tests shrink it)
    float ql_pmax = 0.0f;
    bool ql_mu_ready = false, ql_index_ready = false, ql_stats = false;
    // workspace
    DevBuf offsets, coarse_mat, probe32, hist, start, gstart, groups, order, out_keys, out_cnt, qthr, stats, pq_t, lutf, pq_maxnorm, lutg;
    DevBuf host_xq, host_D, host_I;
    // tensor-core coarse quantizer (K1): split-bf16 centroids, norms, per-call buffers
    DevBuf cent_bf16, cnorm, cmax2, q_bf16, qnorm, cand, cand_score, cand_cnt, flags, nflagged;
    CUtensorMap tmB;
    bool tc_ready = false;
    int coarse_variant = 0;   // 0 = auto (tensor cores when possible), 1 = exact kernels only, 2 = tensor cores with the
                              // full score matrix + radix select (no two-pass filter)
    int kpad = 0;
    // instrumentation
    bool timing = false;
    std::vector<std::array<cudaEvent_t, 8>> evs;   // one event set per query chunk of the last search ([6], [7]: filter kernel)
    cudaEvent_t* ev = nullptr;                     // event set of the chunk being enqueued
    int timed_chunks = 0;
    bool stage_valid = false;
    bool filter_timed = false;   // the last timed search ran the streaming filter kernel (events [6], [7])
    cudaStream_t last_stream = nullptr;
    // small-batch latency path: the whole host-buffer search (H2D, kernels, D2H) replayed as a CUDA graph
    struct GraphEntry {
        cudaGraphExec_t exec = nullptr;
        uint64_t epoch = 0;
        int seen = 0;
    };
    std::map<std::tuple<int64_t, int, int>, GraphEntry> graphs;
    cudaStream_t gstream = nullptr;
    void* pin_xq = nullptr;
    void* pin_D = nullptr;
    void* pin_I = nullptr;
    size_t pin_xq_cap = 0, pin_out_cap = 0;
    uint64_t state_epoch = 0;   // bumped by set_codebooks / set_lists
    int use_graph = 1;          // B200_IVFPQ_GRAPH=0 disables
};

namespace {

int ensure_events(b200_ivfpq_index* h, size_t nchunks) {
    while (h->evs.size() < nchunks) {
        std::array<cudaEvent_t, 8> set{};
        for (auto& e : set) CUDA_TRY(cudaEventCreate(&e));
        h->evs.push_back(set);
    }
    return 0;
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
}

int grid1d(int64_t n, int threads) { return static_cast<int>((n + threads - 1) / threads); }

// K1b: k smallest (distance, id) of every row of a (nq, n) matrix -- radix select, one CTA per row
int launch_select(b200_ivfpq_index* h, const float* mat, int64_t nq, int64_t n, int k, int32_t* ids32, int64_t* ids64,
                  float* dis, cudaStream_t st) {
    const size_t smem = select_smem_bytes(k);
    if (nq < 2 * (int64_t)h->num_sms) {
        select_radix_kernel<1024><<<(unsigned)nq, 1024, smem, st>>>(mat, n, n, k, ids32, ids64, dis);
    } else {
        select_radix_kernel<256><<<(unsigned)nq, 256, smem, st>>>(mat, n, n, k, ids32, ids64, dis);
    }
    LAUNCH_CHECK();
    return 0;
}

int tc_candidates(const b200_ivfpq_index* h, int nprobe) {
    int64_t L = std::max<int64_t>(2 * (int64_t)nprobe, (int64_t)nprobe + 32);
    return (int)std::min<int64_t>(L, 2048);
}

bool tc_usable(const b200_ivfpq_index* h, int nprobe) {
    if (h->coarse_variant == 1 || !h->tc_ready) return false;
    if (nprobe + 32 > 2048) return false;
    return tc_candidates(h, nprobe) < h->nlist;   // otherwise every centroid would be a candidate anyway
}

// one pass of the tensor-core GEMM over a chunk of queries (split-bf16 operand prepared by tc_prepare_queries)
template <int MODE>
int launch_tc_gemm(b200_ivfpq_index* h, int64_t nq, TcGemmParams gp, cudaStream_t st) {
    CUtensorMap tmA;
    if (!tc_make_map(&tmA, h->q_bf16.p, nq, h->kpad, kTcBM))
        return fail(B200_IVFPQ_ECUDA, "cuTensorMapEncodeTiled failed for the query operand");
    gp.cnorm = h->cnorm.as<float>();
    gp.qnorm = h->qnorm.as<float>();
    gp.nq = nq;
    gp.nlist = h->nlist;
    gp.kblocks = h->kpad / kTcBK;
    gp.mtiles = (int)((nq + kTcBM - 1) / kTcBM);
    gp.ntiles = (int)((h->nlist + kTcBN - 1) / kTcBN);
    CUDA_TRY(cudaFuncSetAttribute(coarse_tc_gemm_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)kTcSmemBytes));
    int grid = std::min<int64_t>((int64_t)gp.mtiles * gp.ntiles, h->num_sms);
    coarse_tc_gemm_kernel<MODE><<<grid, kTcThreads, kTcSmemBytes, st>>>(tmA, h->tmB, gp);
    LAUNCH_CHECK();
    return 0;
}

int tc_prepare_queries(b200_ivfpq_index* h, int64_t nq, const float* d_xq, cudaStream_t st) {
    int rc;
    if ((rc = h->q_bf16.ensure(sizeof(__nv_bfloat16) * nq * h->kpad))) return rc;
    if ((rc = h->qnorm.ensure(sizeof(float) * nq))) return rc;
    tc_split_rows_kernel<<<(unsigned)nq, 128, 0, st>>>(d_xq, nq, h->d, h->kpad, 1, h->q_bf16.as<__nv_bfloat16>(),
                                                      h->qnorm.as<float>());
    LAUNCH_CHECK();
    return 0;
}

// approximate scores s(q, c) = ||c||^2 - 2 q.c for one chunk of queries, on the tensor cores
int run_tc_scores(b200_ivfpq_index* h, int64_t nq, const float* d_xq, float* d_scores, cudaStream_t st) {
    int rc;
    if ((rc = tc_prepare_queries(h, nq, d_xq, st))) return rc;
    TcGemmParams gp{};
    gp.out = d_scores;
    return launch_tc_gemm<kTcScores>(h, nq, gp, st);
}

// K1 for one chunk of queries: distances to all centroids, then nprobe-select.
int run_coarse(b200_ivfpq_index* h, int64_t nq, const float* d_xq, int nprobe, int32_t* probe32, int64_t* ids64,
               float* dis, cudaStream_t st, bool time_stages) {
    int rc;
    if (h->coarse_variant == 0 && coarse_small_usable(nq, h->nlist, nprobe)) {
        // latency path: exact distances + register top-32 per 128 centroids, then one select CTA per query
        const int64_t nctas = (h->nlist + kCsThreads - 1) / kCsThreads;
        if ((rc = h->coarse_mat.ensure(sizeof(uint64_t) * nq * nctas * kCsKeep))) return rc;
        const size_t dsm = coarse_small_dist_smem(h->d);
        if ((rc = set_smem(coarse_small_dist_kernel, dsm))) return rc;
        dim3 grid((unsigned)nctas, (unsigned)((nq + kCsQ - 1) / kCsQ));
        coarse_small_dist_kernel<<<grid, kCsThreads, dsm, st>>>(d_xq, h->cent, (int)nq, h->nlist, h->d,
                                                               h->coarse_mat.as<uint64_t>());
        LAUNCH_CHECK();
        if (time_stages) CUDA_TRY(cudaEventRecord(h->ev[1], st));
        const int nkeys = (int)(nctas * kCsKeep);
        const size_t ssm = sizeof(uint64_t) * std::max(nkeys, kThreads);
        if ((rc = set_smem(coarse_small_select_kernel, ssm))) return rc;
        coarse_small_select_kernel<<<(unsigned)nq, kThreads, ssm, st>>>(h->coarse_mat.as<uint64_t>(), nkeys, nprobe,
                                                                       probe32, ids64, dis);
        LAUNCH_CHECK();
        return 0;
    }
    if (tc_usable(h, nprobe)) {
        // tensor-core pre-filter + exact rescoring (coarse_tc.cuh)
        const int L = tc_candidates(h, nprobe);
        const int64_t nchunks = (h->nlist + 31) / 32;
        const bool two_pass = nchunks >= L && h->coarse_variant != 2;
        const int cap = two_pass ? std::min(4 * L, 2048) : L;
        if ((rc = h->cand.ensure(sizeof(int32_t) * nq * std::max(cap, L)))) return rc;
        if ((rc = h->cand_score.ensure(sizeof(float) * nq * L))) return rc;
        if ((rc = h->flags.ensure(sizeof(int) * nq))) return rc;
        if (!h->nflagged.p) {
            if ((rc = h->nflagged.ensure(sizeof(int)))) return rc;
            CUDA_TRY(cudaMemsetAsync(h->nflagged.p, 0, sizeof(int), st));
        }
        const int* cand_cnt = nullptr;
        if (two_pass) {
            // no (nq, nlist) matrix: chunk minima -> tau = L-th smallest minimum -> the same GEMM again as a filter
            if ((rc = h->coarse_mat.ensure(sizeof(float) * nq * nchunks))) return rc;
            if ((rc = h->cand_cnt.ensure(sizeof(int) * nq))) return rc;
            if ((rc = tc_prepare_queries(h, nq, d_xq, st))) return rc;
            TcGemmParams gp{};
            gp.out = h->coarse_mat.as<float>();
            gp.nchunks = nchunks;
            if ((rc = launch_tc_gemm<kTcMinima>(h, nq, gp, st))) return rc;
            if (time_stages) CUDA_TRY(cudaEventRecord(h->ev[1], st));
            if ((rc = launch_select(h, h->coarse_mat.as<float>(), nq, nchunks, L, h->cand.as<int32_t>(), nullptr,
                                    h->cand_score.as<float>(), st)))
                return rc;
            CUDA_TRY(cudaMemsetAsync(h->cand_cnt.p, 0, sizeof(int) * nq, st));
            gp.tau = h->cand_score.as<float>() + (L - 1);
            gp.tau_stride = L;
            gp.cand = h->cand.as<int32_t>();
            gp.cand_cnt = h->cand_cnt.as<int>();
            gp.cap = cap;
            if ((rc = launch_tc_gemm<kTcFilter>(h, nq, gp, st))) return rc;
            cand_cnt = h->cand_cnt.as<int>();
        } else {
            if ((rc = h->coarse_mat.ensure(sizeof(float) * nq * h->nlist))) return rc;
            if ((rc = run_tc_scores(h, nq, d_xq, h->coarse_mat.as<float>(), st))) return rc;
            if (time_stages) CUDA_TRY(cudaEventRecord(h->ev[1], st));
            if ((rc = launch_select(h, h->coarse_mat.as<float>(), nq, h->nlist, L, h->cand.as<int32_t>(), nullptr,
                                    h->cand_score.as<float>(), st)))
                return rc;
        }
        const float eps_rel = 6e-5f, eps_abs = (float)(h->d + 2) * 1.1920929e-7f;
        size_t rsm = rescore_smem_bytes(h->d, nprobe);
        if ((rc = set_smem(coarse_rescore_kernel, rsm))) return rc;
        coarse_rescore_kernel<<<(unsigned)nq, kRescoreThreads, rsm, st>>>(
            d_xq, h->cent, h->qnorm.as<float>(), h->cmax2.as<float>(), h->cand.as<int32_t>(), cand_cnt, cap,
            h->cand_score.as<float>() + (L - 1), L, h->d, h->nlist, nprobe, eps_rel, eps_abs, probe32, ids64, dis,
            h->flags.as<int>());
        LAUNCH_CHECK();
        size_t fsm = sizeof(float) * ((h->d + 3) & ~3) + TopK::smem_bytes(nprobe, kSelCap);
        if ((rc = set_smem(coarse_exact_flagged_kernel, fsm))) return rc;
        coarse_exact_flagged_kernel<<<(unsigned)nq, kThreads, fsm, st>>>(d_xq, h->cent, h->flags.as<int>(), h->d,
                                                                        h->nlist, nprobe, probe32, ids64, dis,
                                                                        h->nflagged.as<int>());
        LAUNCH_CHECK();
        return 0;
    }
    if ((rc = h->coarse_mat.ensure(sizeof(float) * nq * h->nlist))) return rc;
    dim3 grid(static_cast<unsigned>((h->nlist + kCoarseTile - 1) / kCoarseTile),
              static_cast<unsigned>((nq + kCoarseTile - 1) / kCoarseTile));
    coarse_dist_kernel<<<grid, kThreads, 0, st>>>(d_xq, h->cent, h->coarse_mat.as<float>(), (int)nq, h->nlist, h->d,
                                                  h->nlist);
    LAUNCH_CHECK();
    if (time_stages) CUDA_TRY(cudaEventRecord(h->ev[1], st));
    return launch_select(h, h->coarse_mat.as<float>(), nq, h->nlist, nprobe, probe32, ids64, dis, st);
}

int64_t coarse_chunk(const b200_ivfpq_index* h) {
    int64_t qb = static_cast<int64_t>(kCoarseMatrixBudget / (sizeof(float) * h->nlist));
    qb = (qb / kCoarseTile) * kCoarseTile;
    return qb < kCoarseTile ? kCoarseTile : qb;
}

template <int VEC>
int launch_scan(b200_ivfpq_index* h, const ScanParams& sp, int64_t npairs, cudaStream_t st) {
    size_t smem = scan_smem_bytes(sp.M, sp.d, sp.k);
    int rc = set_smem(scan_pairs_kernel<VEC>, smem);
    if (rc) return rc;
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scan_pairs_kernel<VEC>, kThreads, smem));
    if (per_sm < 1) return fail(B200_IVFPQ_EUNSUPPORTED, "scan kernel does not fit (smem %zu B)", smem);
    int64_t grid = static_cast<int64_t>(per_sm) * h->num_sms;
    if (grid > npairs * sp.nseg) grid = npairs * sp.nseg;
    scan_pairs_kernel<VEC><<<(unsigned)grid, kThreads, smem, st>>>(sp);
    LAUNCH_CHECK();
    return 0;
}

// per-index data of the per-query-table filter scan: mu at set_codebooks time, the per-vector term when first needed
int ql_prepare(b200_ivfpq_index* h, cudaStream_t st) {
    int rc;
    if (!h->ql_mu_ready) {
        if ((rc = h->ql_mu.ensure(sizeof(float) * h->d))) return rc;
        if (ql_build_mean(h->cent, h->nlist, h->d, h->ql_mu.as<float>(), st)) return fail(B200_IVFPQ_ECUDA, "ql_build_mean launch failed");
        g_launches.fetch_add(1);
        std::vector<float> mx(h->M);
        CUDA_TRY(cudaStreamSynchronize(st));
        CUDA_TRY(cudaMemcpy(mx.data(), h->pq_maxnorm.p, sizeof(float) * h->M, cudaMemcpyDeviceToHost));
        double a = 0.0;
        for (float v : mx) a += (double)v * (double)v;
        h->ql_pmax = (float)(std::sqrt(a) * 1.0001);
        h->ql_mu_ready = true;
        h->ql_index_ready = false;
    }
    if (!h->ql_index_ready) {
        if ((rc = h->ql_snorm.ensure(sizeof(uint16_t) * std::max<int64_t>(h->ntotal, 1)))) return rc;
        if ((rc = h->ql_sbmin.ensure(sizeof(float) * h->nlist))) return rc;
        if ((rc = h->ql_sbstep.ensure(sizeof(float) * h->nlist))) return rc;
        if (ql_build_index_data(h->cent, h->pq, h->ql_mu.as<float>(), h->offsets.as<int64_t>(), h->codes, h->nlist, h->d,
                                h->M, h->dsub, h->ql_snorm.as<uint16_t>(), h->ql_sbmin.as<float>(),
                                h->ql_sbstep.as<float>(), h->num_sms, st))
            return fail(B200_IVFPQ_ECUDA, "ql_build_index_data launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        g_launches.fetch_add(1);
        h->ql_index_ready = true;
    }
    return 0;
}

// split = nullptr: the whole search.  split != nullptr (b200_ivfpq_search_preassigned_begin): everything up to the
// bootstrap thresholds of the queries [split->lo, split->hi); the rest is enqueued by search_finish.
struct SplitArgs {
    int64_t lo, hi;
    uint32_t* d_thr_out;   // (nq) the thresholds of [lo, hi) are written here, +inf bits elsewhere
};

int search_impl(b200_ivfpq_index* h, int64_t nq, const float* d_xq, int k, int nprobe, const int64_t* d_list_ids,
                float* d_D, int64_t* d_I, cudaStream_t st, const SplitArgs* split = nullptr) {
    if (!h) return fail(B200_IVFPQ_EINVAL, "null index handle");
    if (!h->cent || !h->pq) return fail(B200_IVFPQ_ESTATE, "index is not trained (set_codebooks not called)");
    if (!h->has_lists) return fail(B200_IVFPQ_ESTATE, "inverted lists not set (set_lists not called)");
    if (nq < 0) return fail(B200_IVFPQ_EINVAL, "nq = %lld < 0", (long long)nq);
    if (k < 1 || k > B200_IVFPQ_MAX_K) return fail(B200_IVFPQ_EINVAL, "k = %d out of [1, %d]", k, B200_IVFPQ_MAX_K);
    if (nprobe < 1 || nprobe > B200_IVFPQ_MAX_NPROBE)
        return fail(B200_IVFPQ_EINVAL, "nprobe = %d out of [1, %d]", nprobe, B200_IVFPQ_MAX_NPROBE);
    if (nq == 0) return 0;
    if (!d_xq || (!split && (!d_D || !d_I))) return fail(B200_IVFPQ_EINVAL, "null query / result pointer");
    CUDA_TRY(cudaSetDevice(h->device));
    h->pend.valid = false;
    if (!d_list_ids && nprobe > h->nlist) nprobe = static_cast<int>(h->nlist);   // Faiss clamps nprobe to nlist

    const bool legacy = h->scan_variant == 6;   // B200_IVFPQ_SCAN=legacy: round 1's kernel selection (no scan_qlut)
    const int sv = legacy ? 0 : h->scan_variant;
    const bool timing = h->timing;
    h->stage_valid = false;
    h->filter_timed = false;
    h->timed_chunks = 0;
    h->last_stream = st;

    // query chunking keeps the workspace bounded
    int64_t qb = nq;
    if (!d_list_ids) qb = std::min<int64_t>(qb, coarse_chunk(h));
    qb = std::min<int64_t>(qb, std::max<int64_t>(1, (int64_t)(kPairOutBudget / (sizeof(uint64_t) * (size_t)nprobe * k))));
    qb = std::min<int64_t>(qb, (int64_t)((1ll << 30) / nprobe));
    if (split && qb < nq) return fail(B200_IVFPQ_EUNSUPPORTED, "split search: the batch does not fit one query chunk");
    if (timing) {
        int rc0 = ensure_events(h, (size_t)((nq + qb - 1) / qb));
        if (rc0) return rc0;
    }

    // small batches: split every (query, probe) pair into nseg list segments so that the scan fills the GPU
    int nseg = 1;
    {
        const int64_t pairs = std::min<int64_t>(qb, nq) * nprobe;
        const int64_t target = 2 * (int64_t)h->num_sms;
        // as many segments as keep all work items in ONE wave of resident CTAs (two per SM); at least two below one
        // wave's worth of pairs (shorter critical path per CTA, the LUTs are prebuilt anyway)
        if (pairs < target) nseg = (int)std::max<int64_t>(2, std::min<int64_t>(16, target / pairs));
        if (sv >= 3) nseg = 1;   // forced multi-query kernels (tests): they scan whole lists
        if (h->force_nseg > 0) nseg = h->force_nseg;
    }

    int rc;
    if ((rc = h->stats.ensure(sizeof(PairStats)))) return rc;
    if ((rc = h->probe32.ensure(sizeof(int32_t) * qb * nprobe))) return rc;
    if ((rc = h->hist.ensure(sizeof(int) * h->nlist * kQlHostBuckets))) return rc;     // scan_qlut sorts by (rank bucket, list)
    if ((rc = h->start.ensure(sizeof(int) * h->nlist * kQlHostBuckets))) return rc;
    if ((rc = h->gstart.ensure(sizeof(int) * h->nlist * kQlHostBuckets))) return rc;
    if ((rc = h->groups.ensure(sizeof(QuadGroup) * qb * nprobe))) return rc;
    if ((rc = h->order.ensure(sizeof(int32_t) * qb * nprobe))) return rc;
    if ((rc = h->out_keys.ensure(sizeof(uint64_t) * qb * nprobe * k * nseg))) return rc;
    if ((rc = h->out_cnt.ensure(sizeof(int) * qb * nprobe * nseg))) return rc;
    if ((rc = h->qthr.ensure(sizeof(uint32_t) * qb))) return rc;
    CUDA_TRY(cudaMemsetAsync(h->stats.p, 0, sizeof(PairStats), st));

    for (int64_t q0 = 0; q0 < nq; q0 += qb) {
        const int64_t nqc = std::min<int64_t>(qb, nq - q0);
        const int64_t npairs = nqc * nprobe;
        const float* xq = d_xq + q0 * h->d;
        const bool tm = timing;
        if (tm) h->ev = h->evs[h->timed_chunks].data();
        int32_t* probe32 = h->probe32.as<int32_t>();

        // which scan kernel will run decides how pairs are grouped (2 or 4 queries of a list per work item)
        const bool aligned16 = (reinterpret_cast<uintptr_t>(h->codes) & 15) == 0;
        int quad_ctas = 0;
        if (quad_supported(h->M, h->d, k) && aligned16 && nseg == 1 && h->max_list < (int64_t)kQuadMaxList &&
            (sv == 4 ||
             (sv == 0 && npairs >= 6 * h->nlist && h->ntotal >= kQuadMinAvgList * std::max<int64_t>(h->nonempty, 1))))
            quad_ctas = quad_grid(h->M, h->dsub, h->d, k, npairs, h->num_sms);
        if (sv == 4 && quad_ctas == 0)
            return fail(B200_IVFPQ_EUNSUPPORTED, "four-query scan kernel unsupported for M=%d d=%d k=%d", h->M, h->d, k);
        // the per-query-table filter (scan_qlut.cuh) replaces both multi-query kernels wherever lists are shared by
        // queries: its per-work-item cost is a table copy instead of a table build.  For M = 32 / 64 it also replaces
        // the lane-per-code generic kernel (bank-conflicted 32 / 64 KB tables) at any number of queries per list
        int ql_ctas = 0;
        if (ql_supported_host(h->M, h->d, k) && aligned16 && nseg == 1 && h->max_list < (int64_t)kQlHostMaxList &&
            (sv == 5 || (sv == 0 && !legacy && (2 * npairs >= h->nlist || h->M != 16))))
            ql_ctas = ql_grid(h->M, h->d, k, npairs, h->num_sms);
        if (sv == 5 && ql_ctas == 0)
            return fail(B200_IVFPQ_EUNSUPPORTED, "per-query-table scan kernel unsupported for M=%d d=%d k=%d", h->M, h->d, k);
        int st_ctas = 0;
        StHostBuffers sb{};
        if (ql_ctas && h->st_mode) st_ctas = st_filter_grid(h->M, npairs, h->num_sms);
        // Lists probed by one or two queries (at most one pair per list on average): work items of two pairs and 32-bit
        // table words -- half the shared-memory wavefronts per code (st_filter_kernel<16, true>).
        // B200_IVFPQ_STREAM_TWO: 0 never, 1 always, 2 = the bulk-async experiment (scan_stream.cuh A2), default auto
        bool two = false;
        sb.two_kind = 0;
        if (st_ctas && h->M == 16) {
            if (h->st_two == 2 && !split) {
                const int g2 = st_filter2_grid(npairs, h->num_sms);
                if (g2 > 0) {
                    two = true;
                    sb.two_kind = 2;
                    st_ctas = g2;
                }
            } else if (h->st_two == 1 || (h->st_two < 0 && npairs <= h->nlist)) {
                two = true;
                sb.two_kind = 1;
            }
        }
        sb.gsz = two ? 2 : 4;
        if (st_ctas) {
            // survivor records: a share of the (query, code) pairs the batch scans; slabs: keys at or below a
            // query's bootstrap threshold.  An overflow is detected on the device and answered by the fallback launches.
            const double visits = (double)npairs * (double)h->ntotal / (double)std::max<int64_t>(h->nonempty, 1);
            const double want = std::min(std::max(visits * h->st_rate, h->st_minrec), 192.0 * 1048576.0);
            sb.max_chunks = (unsigned int)(want / kStChunkRecords);
            // slabs: room for thousands of duplicate codes at a query's threshold distance (thresholds that came from
            // another shard admit by distance only), within 2 GB per query chunk
            sb.capq = h->st_capq > 0 ? h->st_capq
                                     : (int)std::max<int64_t>(std::max(4096, 32 * k),
                                                              std::min<int64_t>(16384, (int64_t(2) << 30) / (8 * std::max<int64_t>(qb, 1))));
            if ((rc = h->st_srec.ensure((size_t)sb.max_chunks * kStChunkRecords * 8))) return rc;
            if ((rc = h->st_sfill.ensure((size_t)sb.max_chunks * 4))) return rc;
            if ((rc = h->st_ctr.ensure(kStCtrBytes))) return rc;
            if ((rc = h->st_slab.ensure((size_t)qb * sb.capq * 8))) return rc;
            if ((rc = h->st_qcnt.ensure((size_t)qb * 4))) return rc;
            if ((rc = h->st_qflag.ensure((size_t)qb * 4))) return rc;
            if ((rc = h->st_qkey.ensure((size_t)qb * 8))) return rc;
            if ((rc = h->st_prefix.ensure((size_t)qb * nprobe * 4))) return rc;
            if ((rc = h->st_pdis.ensure((size_t)qb * nprobe * 4))) return rc;
            sb.srec = h->st_srec.p;
            sb.sfill = h->st_sfill.p;
            sb.ctr = h->st_ctr.p;
            sb.slab = h->st_slab.p;
            sb.qcnt = h->st_qcnt.p;
            sb.qflag = h->st_qflag.p;
            sb.qkey = h->st_qkey.p;
            sb.prefix = h->st_prefix.p;
            sb.pdis = h->st_pdis.p;
        }
        if (ql_ctas) {
            quad_ctas = 0;
            if ((rc = ql_prepare(h, st))) return rc;
            // (tables prepared ahead of this call live in these buffers: a reallocation drops them)
            if (h->prep.valid && (h->ql_lut.cap < sizeof(uint16_t) * (size_t)qb * 256 * h->M || h->ql_scale.cap < sizeof(float) * qb ||
                                  h->ql_amin.cap < sizeof(float) * qb))
                h->prep.valid = false;
            if ((rc = h->ql_lut.ensure(sizeof(uint16_t) * (size_t)qb * 256 * h->M))) return rc;
            if ((rc = h->ql_scale.ensure(sizeof(float) * qb))) return rc;
            if ((rc = h->ql_amin.ensure(sizeof(float) * qb))) return rc;
            if (h->ql_stats && !h->ql_counters.p) {
                if ((rc = h->ql_counters.ensure(3 * sizeof(unsigned long long)))) return rc;
                CUDA_TRY(cudaMemsetAsync(h->ql_counters.p, 0, 3 * sizeof(unsigned long long), st));
            }
        }
        if (split && !ql_ctas) return fail(B200_IVFPQ_EUNSUPPORTED, "split search: the per-query-table scan does not apply here");
        const int gsz = two ? 2 : (quad_ctas || ql_ctas) ? 4 : 2;
        if (quad_ctas && (rc = h->lutf.ensure(sizeof(float4) * quad_scratch_float4(h->M) * quad_ctas))) return rc;

        // the per-query tables only need the queries: they are built on a side stream while the coarse quantizer and the
        // pair set-up run (not under stream capture: the latency path's CUDA graph keeps one stream)
        bool tables_async = false;
        if (ql_ctas && h->prep.valid && h->prep.xq == xq && h->prep.nq == nqc && nqc == nq) {
            tables_async = true;      // b200_ivfpq_prepare_queries started them: ev_join is recorded
        } else if (ql_ctas) {
            cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
            CUDA_TRY(cudaStreamIsCapturing(st, &cs));
            if (cs == cudaStreamCaptureStatusNone && h->side_ok()) {
                CUDA_TRY(cudaEventRecord(h->ev_fork, st));
                CUDA_TRY(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
                if (ql_build_query_tables(xq, nqc, h->pq, h->ql_mu.as<float>(), h->pq_maxnorm.as<float>(), h->d, h->M, h->dsub,
                                          h->ql_lut.as<uint16_t>(), h->ql_scale.as<float>(), h->ql_amin.as<float>(), h->side))
                    return fail(B200_IVFPQ_ECUDA, "query table launch failed: %s", cudaGetErrorString(cudaGetLastError()));
                g_launches.fetch_add(1);
                CUDA_TRY(cudaEventRecord(h->ev_join, h->side));
                tables_async = true;
            }
        }

        h->prep.valid = false;        // one search per prepare
        if (tm) CUDA_TRY(cudaEventRecord(h->ev[0], st));
        if (d_list_ids) {
            probes_from_i64_kernel<<<grid1d(npairs, 256), 256, 0, st>>>(d_list_ids + q0 * nprobe, probe32, npairs,
                                                                       h->nlist);
            LAUNCH_CHECK();
            if (tm) CUDA_TRY(cudaEventRecord(h->ev[1], st));
        } else {
            if ((rc = run_coarse(h, nqc, xq, nprobe, probe32, nullptr, nullptr, st, tm))) return rc;
        }
        if (tm) CUDA_TRY(cudaEventRecord(h->ev[2], st));

        // pair setup
        PairStats* stats = h->stats.as<PairStats>();
        const bool small_setup = nseg > 1 && npairs <= 8192 && npairs * nseg <= (1 << 17) && !quad_ctas && !ql_ctas;
        if (small_setup) {
            // latency path: one kernel instead of two memsets and four kernels; pairs keep their natural order
            pair_setup_small_kernel<<<1, 1024, 0, st>>>(probe32, (int)npairs, h->offsets.as<int64_t>(),
                                                        h->order.as<int32_t>(), h->qthr.as<uint32_t>(), (int)nqc,
                                                        h->out_cnt.as<int>(), (int)(npairs * nseg), stats);
            LAUNCH_CHECK();
        } else {
            // (rank bucket, list) keys: the in-kernel path tightens thresholds as it goes and wants every query's nearest
            // list first; the streaming filter's thresholds are fixed, it only needs "same list together"
            const int nbuckets = (ql_ctas && !st_ctas) ? kQlHostBuckets : 1;
            const int64_t nkeys = h->nlist * nbuckets;
            CUDA_TRY(cudaMemsetAsync(h->hist.p, 0, sizeof(int) * nkeys, st));
            CUDA_TRY(cudaMemsetAsync(h->out_cnt.p, 0, sizeof(int) * npairs * nseg, st));
            fill_u32_kernel<<<grid1d(nqc, 256), 256, 0, st>>>(h->qthr.as<uint32_t>(), nqc, kInfBits);
            LAUNCH_CHECK();
            if (ql_ctas) {
                if (ql_launch_hist(probe32, npairs, nprobe, h->nlist, nbuckets, h->offsets.as<int64_t>(), h->hist.as<int>(), stats, st))
                    return fail(B200_IVFPQ_ECUDA, "pair histogram launch failed");
                g_launches.fetch_add(1);
            } else {
                pair_hist_kernel<<<grid1d(npairs, 256), 256, 0, st>>>(probe32, npairs, h->offsets.as<int64_t>(),
                                                                     h->hist.as<int>(), stats);
                LAUNCH_CHECK();
            }
            pair_scan_kernel<<<1, 1024, 0, st>>>(h->hist.as<int>(), h->start.as<int>(), h->gstart.as<int>(), nkeys, gsz,
                                                 stats);
            LAUNCH_CHECK();
            CUDA_TRY(cudaMemsetAsync(h->groups.p, 0xff, (ql_ctas ? kQlGroupBytes : gsz == 4 ? sizeof(QuadGroup) : sizeof(DuoGroup)) * npairs, st));
            if (ql_ctas) {
                if (ql_launch_scatter(probe32, npairs, nprobe, h->nlist, nbuckets, h->offsets.as<int64_t>(), h->start.as<int>(),
                                      h->gstart.as<int>(), h->hist.as<int>(), h->order.as<int32_t>(), h->groups.p, gsz, st))
                    return fail(B200_IVFPQ_ECUDA, "pair scatter launch failed");
                g_launches.fetch_add(1);
                // one table per QUERY (not per pair): A_q quantised with the query's own scale
                if (tables_async) {
                    CUDA_TRY(cudaStreamWaitEvent(st, h->ev_join, 0));
                } else {
                    if (ql_build_query_tables(xq, nqc, h->pq, h->ql_mu.as<float>(), h->pq_maxnorm.as<float>(), h->d, h->M,
                                              h->dsub, h->ql_lut.as<uint16_t>(), h->ql_scale.as<float>(),
                                              h->ql_amin.as<float>(), st))
                        return fail(B200_IVFPQ_ECUDA, "query table launch failed: %s", cudaGetErrorString(cudaGetLastError()));
                    g_launches.fetch_add(1);
                }
            } else {
                pair_scatter_kernel<<<grid1d(npairs, 256), 256, 0, st>>>(probe32, npairs, h->offsets.as<int64_t>(),
                                                                        h->start.as<int>(), h->gstart.as<int>(),
                                                                        h->hist.as<int>(), h->order.as<int32_t>(),
                                                                        h->groups.p, gsz);
                LAUNCH_CHECK();
            }
        }
        if (tm) CUDA_TRY(cudaEventRecord(h->ev[3], st));

        // K2+K3+K4
        const int* stream_guard = nullptr;   // set when the streaming pipeline ran: what follows is its fallback
        const int* stream_qflag = nullptr;
        ScanParams sp;
        sp.xq = xq;
        sp.cent = h->cent;
        sp.pq = h->pq;
        sp.offsets = h->offsets.as<int64_t>();
        sp.codes = h->codes;
        sp.probe = probe32;
        sp.order = h->order.as<int32_t>();
        sp.groups = h->groups.p;
        sp.lutf_scratch = h->lutf.as<float4>();
        sp.pq_maxnorm = h->pq_maxnorm.as<float>();
        sp.out_keys = h->out_keys.as<uint64_t>();
        sp.out_cnt = h->out_cnt.as<int>();
        sp.qthr = h->qthr.as<uint32_t>();
        sp.stats = stats;
        sp.d = h->d;
        sp.M = h->M;
        sp.dsub = h->dsub;
        sp.nprobe = nprobe;
        sp.k = k;
        sp.nseg = nseg;
        sp.lutg = nullptr;
        sp.negzero2 = 0x8000000080000000ull;
        sp.quad_drain_at = h->quad_drain_at;
        bool use_skew = skew_supported(h->M, h->d, k) && aligned16 && sv != 1;
        if (nseg > 1 && npairs <= 1024) {
            // small batches: every pair is scanned by nseg CTAs -- build its LUT once instead of nseg times
            if ((rc = h->lutg.ensure(sizeof(float) * npairs * h->M * 256))) return rc;
            dim3 lgrid((unsigned)npairs, (unsigned)((h->M + 7) / 8));
            lut_small_kernel<<<lgrid, 256, 0, st>>>(xq, h->cent, h->pq, probe32, nprobe, h->d, h->M, h->dsub,
                                                   use_skew ? 1 : 0, h->lutg.as<float>());
            LAUNCH_CHECK();
            sp.lutg = h->lutg.as<float>();
        }
        if (sv >= 2 && sv != 4 && sv != 5 && !use_skew &&
            !(sv == 3 && duo32_supported(h->M, h->d, k)))
            return fail(B200_IVFPQ_EUNSUPPORTED, "skewed scan kernel unsupported for M=%d d=%d k=%d", h->M, h->d, k);
        // two queries per work item pay off once lists are shared.  With Poisson-distributed queries per list, one
        // probing query per list on average already fills 70 % of the slots (4.8 x 0.7 > the one-query kernel's
        // 2.9 TB/s); for M = 32, whose only alternative is the bank-conflicted generic kernel (2.1 TB/s), even
        // half-empty work items win (C5 sweep: profiles/r1_sweep_c5_batch_nprobe.json)
        bool use_duo = use_skew && nseg == 1 && sv != 2 &&
                       (sv == 3 || npairs >= h->nlist);
        if (ql_ctas) {
            QlHostParams qp;
            qp.snorm = h->ql_snorm.as<uint16_t>();
            qp.sbmin = h->ql_sbmin.as<float>();
            qp.sbstep = h->ql_sbstep.as<float>();
            qp.pmax = h->ql_pmax;
            qp.qlut = h->ql_lut.as<uint16_t>();
            qp.qscale = h->ql_scale.as<float>();
            qp.qamin = h->ql_amin.as<float>();
            qp.counters = h->ql_stats ? h->ql_counters.as<unsigned long long>() : nullptr;
            qp.guard = nullptr;
            qp.qflag = nullptr;
            if (split) {
                if (!st_ctas) return fail(B200_IVFPQ_EUNSUPPORTED, "split search: the streaming pipeline does not apply here");
                if (st_launch_boot(sp, qp, sb, nqc, split->lo, split->hi, st))
                    return fail(B200_IVFPQ_ECUDA, "bootstrap launch failed: %s", cudaGetErrorString(cudaGetLastError()));
                g_launches.fetch_add(1);
                CUDA_TRY(cudaMemcpyAsync(split->d_thr_out, sp.qthr, sizeof(uint32_t) * nqc, cudaMemcpyDeviceToDevice, st));
                h->pend.valid = true;
                h->pend.sp = sp;
                h->pend.qp = qp;
                h->pend.sb = sb;
                h->pend.nq = nqc;
                h->pend.k = k;
                h->pend.nprobe = nprobe;
                h->pend.st_ctas = st_ctas;
                h->pend.ql_ctas = ql_ctas;
                h->pend.st = st;
                if (tm) {
                    CUDA_TRY(cudaEventRecord(h->ev[4], st));
                    CUDA_TRY(cudaEventRecord(h->ev[5], st));
                }
                return 0;
            }
            if (st_ctas) {
                // streaming pipeline: thresholds -> filter -> exact evaluation -> select; D / I are final after it ...
                if (tm) h->filter_timed = true;
                if (st_launch(sp, qp, sb, nqc, h->ids, d_D + q0 * k, d_I + q0 * k, st_ctas, h->num_sms, st,
                              tm ? h->ev[6] : nullptr, tm ? h->ev[7] : nullptr))
                    return fail(B200_IVFPQ_ECUDA, "streaming scan launch failed: %s", cudaGetErrorString(cudaGetLastError()));
                g_launches.fetch_add(4);
                // ... unless a buffer overflowed: then the guarded launches below recompute the batch
                qp.guard = st_overflow_flag(sb.ctr);
                qp.qflag = h->st_qflag.as<int>();
            }
            if (ql_launch_scan(sp, qp, ql_ctas, st))
                return fail(B200_IVFPQ_ECUDA, "per-query-table scan launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            g_launches.fetch_add(1);
            stream_guard = qp.guard;
            stream_qflag = qp.qflag;
        } else if (quad_ctas) {
            if ((rc = launch_scan_quad(sp, h->pq_t.as<float>(), quad_ctas, st)))
                return fail(B200_IVFPQ_ECUDA, "four-query scan launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            g_launches.fetch_add(1);
        } else if (use_duo) {
            if ((rc = launch_scan_duo(sp, h->pq_t.as<float>(), npairs, h->num_sms, st))) {
                if (rc == -1) return fail(B200_IVFPQ_ECUDA, "two-query scan launch failed: %s",
                                          cudaGetErrorString(cudaGetLastError()));
                return rc;
            }
            g_launches.fetch_add(1);
        } else if (duo32_supported(h->M, h->d, k) && aligned16 && nseg == 1 && sv != 1 &&
                   sv != 2 && (sv == 3 || 2 * npairs >= h->nlist) &&
                   (rc = launch_scan_duo32(sp, h->pq_t.as<float>(), npairs, h->num_sms, st)) != -2) {
            // M = 32: two queries per work item, two alternating tables (scan_duo32.cuh)
            if (rc == -1) return fail(B200_IVFPQ_ECUDA, "two-query M=32 scan launch failed: %s",
                                      cudaGetErrorString(cudaGetLastError()));
            g_launches.fetch_add(1);
        } else if (use_skew) {
            if ((rc = launch_scan_skew(sp, h->pq_t.as<float>(), npairs, h->num_sms, st))) {
                if (rc == -1) return fail(B200_IVFPQ_ECUDA, "skewed scan launch failed: %s",
                                          cudaGetErrorString(cudaGetLastError()));
                return rc;
            }
            g_launches.fetch_add(1);
        } else if (h->M % 16 == 0 && aligned16) {
            if ((rc = launch_scan<16>(h, sp, npairs, st))) return rc;
        } else if (h->M % 8 == 0 && (reinterpret_cast<uintptr_t>(h->codes) & 7) == 0) {
            if ((rc = launch_scan<8>(h, sp, npairs, st))) return rc;
        } else if (h->M % 4 == 0 && (reinterpret_cast<uintptr_t>(h->codes) & 3) == 0) {
            if ((rc = launch_scan<4>(h, sp, npairs, st))) return rc;
        } else {
            if ((rc = launch_scan<1>(h, sp, npairs, st))) return rc;
        }
        if (tm) CUDA_TRY(cudaEventRecord(h->ev[4], st));

        // K4b
        size_t msmem = TopK::smem_bytes(k, kMergeCap);
        if ((rc = set_smem(merge_query_kernel, msmem))) return rc;
        merge_query_kernel<<<(unsigned)nqc, kThreads, msmem, st>>>(h->out_keys.as<uint64_t>(), h->out_cnt.as<int>(),
                                                                  probe32, h->offsets.as<int64_t>(), h->ids, nprobe, k,
                                                                  nseg, h->qthr.as<uint32_t>(), d_D + q0 * k, d_I + q0 * k,
                                                                  stream_guard, stream_qflag);
        LAUNCH_CHECK();
        if (tm) {
            CUDA_TRY(cudaEventRecord(h->ev[5], st));
            h->timed_chunks++;
            h->stage_valid = true;
        }
    }
    return 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
extern "C" {

const char* b200_ivfpq_last_error(void) { return g_last_error.c_str(); }

const char* b200_ivfpq_version(void) { return "b200-ivfpq 0.1 (sm_100a)"; }

int64_t b200_ivfpq_launch_count(void) { return g_launches.load(); }

int b200_ivfpq_create(int d, int64_t nlist, int m, int nbits, b200_ivfpq_t* out) {
    if (!out) return fail(B200_IVFPQ_EINVAL, "out == NULL");
    *out = nullptr;
    if (d <= 0 || nlist <= 0 || m <= 0) return fail(B200_IVFPQ_EINVAL, "d, nlist and m must be positive");
    if (nbits != 8) return fail(B200_IVFPQ_EUNSUPPORTED, "nbits = %d: only 8-bit PQ codes are supported", nbits);
    if (d % m != 0) return fail(B200_IVFPQ_EINVAL, "d = %d is not a multiple of m = %d", d, m);
    if (nlist >= (int64_t(1) << 31)) return fail(B200_IVFPQ_EINVAL, "nlist too large");
    if (m > 256) return fail(B200_IVFPQ_EUNSUPPORTED, "m = %d > 256", m);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(B200_IVFPQ_ECUDA, "no CUDA device available (%s): this library has no CPU path",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    }
    b200_ivfpq_index* h = new b200_ivfpq_index();
    h->d = d;
    h->nlist = nlist;
    h->M = m;
    h->nbits = nbits;
    h->dsub = d / m;
    cudaError_t e2 = cudaGetDevice(&h->device);
    if (e2 == cudaSuccess) e2 = cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, h->device);
    if (e2 != cudaSuccess) {
        delete h;
        return fail(B200_IVFPQ_ECUDA, "cannot query the current CUDA device: %s", cudaGetErrorString(e2));
    }
    const char* v = getenv("B200_IVFPQ_SCAN");
    if (v) h->scan_variant = !strcmp(v, "generic") ? 1 : !strcmp(v, "skew") ? 2 : !strcmp(v, "duo") ? 3 : !strcmp(v, "quad") ? 4 : !strcmp(v, "qlut") ? 5 : !strcmp(v, "legacy") ? 6 : 0;
    v = getenv("B200_IVFPQ_GRAPH");
    if (v) h->use_graph = atoi(v) != 0;
    v = getenv("B200_IVFPQ_NSEG");
    if (v) h->force_nseg = std::max(0, std::min(16, atoi(v)));
    v = getenv("B200_IVFPQ_QUAD_DRAIN");
    if (v) h->quad_drain_at = std::max(0, std::min(256, atoi(v)));
    v = getenv("B200_IVFPQ_STREAM");
    if (v) h->st_mode = atoi(v);
    v = getenv("B200_IVFPQ_STREAM_TWO");
    if (v) h->st_two = atoi(v);
    v = getenv("B200_IVFPQ_STREAM_CAPQ");
    if (v) h->st_capq = std::max(16, atoi(v));
    v = getenv("B200_IVFPQ_STREAM_MINREC");
    if (v) h->st_minrec = std::max(64.0, atof(v));
    v = getenv("B200_IVFPQ_STREAM_RATE");
    if (v) h->st_rate = std::max(1e-6, atof(v));
    v = getenv("B200_IVFPQ_QL_STATS");
    if (v) h->ql_stats = atoi(v) != 0;
    v = getenv("B200_IVFPQ_COARSE");
    if (v) h->coarse_variant = !strcmp(v, "exact") ? 1 : !strcmp(v, "matrix") ? 2 : 0;
    *out = h;
    return 0;
}

int b200_ivfpq_destroy(b200_ivfpq_t h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    DevBuf* bufs[] = {&h->offsets, &h->coarse_mat, &h->probe32, &h->hist,   &h->start,  &h->pq_t, &h->gstart, &h->groups, &h->lutf, &h->pq_maxnorm, &h->lutg,
                      &h->order,   &h->out_keys,   &h->out_cnt, &h->qthr,   &h->stats,  &h->host_xq,
                      &h->host_D,  &h->host_I,     &h->cent_bf16, &h->cnorm, &h->cmax2,
                      &h->q_bf16,  &h->qnorm,      &h->cand,   &h->cand_score, &h->flags, &h->nflagged, &h->cand_cnt,
                      &h->ql_mu,   &h->ql_snorm,   &h->ql_sbmin, &h->ql_sbstep, &h->ql_lut, &h->ql_scale, &h->ql_amin, &h->ql_counters,
                      &h->st_srec, &h->st_sfill, &h->st_ctr, &h->st_slab, &h->st_qcnt, &h->st_qflag, &h->st_qkey, &h->st_prefix, &h->st_pdis};
    for (DevBuf* b : bufs) b->release();
    for (auto& set : h->evs)
        for (auto& e : set)
            if (e) cudaEventDestroy(e);
    for (auto& kv : h->graphs)
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    if (h->gstream) cudaStreamDestroy(h->gstream);
    if (h->side) cudaStreamDestroy(h->side);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->pin_xq) cudaFreeHost(h->pin_xq);
    if (h->pin_D) cudaFreeHost(h->pin_D);
    if (h->pin_I) cudaFreeHost(h->pin_I);
    delete h;
    return 0;
}

int b200_ivfpq_set_codebooks(b200_ivfpq_t h, const float* d_centroids, const float* d_pq) {
    if (!h) return fail(B200_IVFPQ_EINVAL, "null index handle");
    if (!d_centroids) return fail(B200_IVFPQ_EINVAL, "null centroid pointer");
    h->cent = d_centroids;
    h->pq = d_pq;
    h->state_epoch++;
    h->ql_mu_ready = false;
    h->ql_index_ready = false;
    h->prep.valid = false;
    {   // K1 tensor-core operands: B' = [ch | cl | ch] (nlist, kpad) bf16, ||c||^2, max ||c||^2
        CUDA_TRY(cudaSetDevice(h->device));
        h->tc_ready = false;
        h->kpad = tc_kpad(h->d);
        int rc0;
        if ((rc0 = h->cent_bf16.ensure(sizeof(__nv_bfloat16) * h->nlist * h->kpad))) return rc0;
        if ((rc0 = h->cnorm.ensure(sizeof(float) * h->nlist))) return rc0;
        if ((rc0 = h->cmax2.ensure(sizeof(float)))) return rc0;
        tc_split_rows_kernel<<<(unsigned)h->nlist, 128>>>(d_centroids, h->nlist, h->d, h->kpad, 0,
                                                          h->cent_bf16.as<__nv_bfloat16>(), h->cnorm.as<float>());
        LAUNCH_CHECK();
        tc_max_kernel<<<1, 256>>>(h->cnorm.as<float>(), h->nlist, h->cmax2.as<float>());
        LAUNCH_CHECK();
        CUDA_TRY(cudaDeviceSynchronize());
        h->tc_ready = tc_make_map(&h->tmB, h->cent_bf16.p, h->nlist, h->kpad, kTcBN);
    }
    if (!d_pq) return 0;   // coarse-only handle (IndexFlatL2): b200_ivfpq_coarse works, search does not
    // m-fastest copy of the PQ codebook for the conflict-free scan kernel's LUT build
    CUDA_TRY(cudaSetDevice(h->device));
    const int64_t total = static_cast<int64_t>(h->M) * 256 * h->dsub;
    int rc = h->pq_t.ensure(sizeof(float) * total);
    if (rc) return rc;
    pq_transpose_kernel<<<grid1d(total, 256), 256>>>(d_pq, h->pq_t.as<float>(), h->M, h->dsub);
    LAUNCH_CHECK();
    if ((rc = h->pq_maxnorm.ensure(sizeof(float) * h->M))) return rc;
    pq_maxnorm_kernel<<<h->M, 256>>>(d_pq, h->dsub, h->pq_maxnorm.as<float>());   // scan_quad.cuh's quantisation bound
    LAUNCH_CHECK();
    CUDA_TRY(cudaDeviceSynchronize());
    return 0;
}

int b200_ivfpq_set_lists(b200_ivfpq_t h, const int64_t* h_offsets, const uint8_t* d_codes, const int64_t* d_ids,
                         int64_t ntotal) {
    if (!h) return fail(B200_IVFPQ_EINVAL, "null index handle");
    if (!h_offsets) return fail(B200_IVFPQ_EINVAL, "null offsets");
    if (ntotal < 0 || h_offsets[0] != 0 || h_offsets[h->nlist] != ntotal)
        return fail(B200_IVFPQ_EINVAL, "offsets must start at 0 and end at ntotal");
    int64_t max_list = 0, nonempty = 0;
    for (int64_t l = 0; l < h->nlist; l++) {
        int64_t sz = h_offsets[l + 1] - h_offsets[l];
        max_list = std::max(max_list, sz);
        nonempty += sz > 0;
        if (sz < 0) return fail(B200_IVFPQ_EINVAL, "offsets not monotone at list %lld", (long long)l);
        if (sz >= (int64_t(1) << 32)) return fail(B200_IVFPQ_EUNSUPPORTED, "list %lld longer than 2^32", (long long)l);
    }
    if (ntotal > 0 && !d_codes) return fail(B200_IVFPQ_EINVAL, "null codes with ntotal > 0");
    CUDA_TRY(cudaSetDevice(h->device));
    int rc = h->offsets.ensure(sizeof(int64_t) * (h->nlist + 1));
    if (rc) return rc;
    CUDA_TRY(cudaMemcpy(h->offsets.p, h_offsets, sizeof(int64_t) * (h->nlist + 1), cudaMemcpyHostToDevice));
    h->codes = d_codes;
    h->ids = d_ids;
    h->ntotal = ntotal;
    h->max_list = max_list;
    h->nonempty = nonempty;
    h->has_lists = true;
    h->state_epoch++;
    h->ql_index_ready = false;
    return 0;
}

int b200_ivfpq_coarse(b200_ivfpq_t h, int64_t nq, const float* d_xq, int nprobe, int64_t* d_ids, float* d_dis,
                      void* stream) {
    if (!h) return fail(B200_IVFPQ_EINVAL, "null index handle");
    if (!h->cent) return fail(B200_IVFPQ_ESTATE, "index is not trained (set_codebooks not called)");
    if (nprobe < 1 || nprobe > B200_IVFPQ_MAX_NPROBE)
        return fail(B200_IVFPQ_EINVAL, "nprobe = %d out of [1, %d]", nprobe, B200_IVFPQ_MAX_NPROBE);
    if (nq < 0) return fail(B200_IVFPQ_EINVAL, "nq < 0");
    if (nq == 0) return 0;
    if (!d_xq || !d_ids) return fail(B200_IVFPQ_EINVAL, "null pointer");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int64_t qb = coarse_chunk(h);
    for (int64_t q0 = 0; q0 < nq; q0 += qb) {
        int64_t nqc = std::min<int64_t>(qb, nq - q0);
        int rc = run_coarse(h, nqc, d_xq + q0 * h->d, nprobe, nullptr, d_ids + q0 * nprobe,
                            d_dis ? d_dis + q0 * nprobe : nullptr, st, false);
        if (rc) return rc;
    }
    return 0;
}

int b200_ivfpq_coarse_scores(b200_ivfpq_t h, int64_t nq, const float* d_xq, float* d_scores, void* stream) {
    if (!h) return fail(B200_IVFPQ_EINVAL, "null index handle");
    if (!h->cent) return fail(B200_IVFPQ_ESTATE, "index is not trained (set_codebooks not called)");
    if (!h->tc_ready || h->coarse_variant == 1)
        return fail(B200_IVFPQ_EUNSUPPORTED, "tensor-core coarse path is not available / disabled");
    if (nq <= 0 || !d_xq || !d_scores) return fail(B200_IVFPQ_EINVAL, "bad arguments");
    CUDA_TRY(cudaSetDevice(h->device));
    return run_tc_scores(h, nq, d_xq, d_scores, reinterpret_cast<cudaStream_t>(stream));
}

int b200_ivfpq_coarse_fallbacks(b200_ivfpq_t h, int64_t* h_count) {
    if (!h || !h_count) return fail(B200_IVFPQ_EINVAL, "null pointer");
    *h_count = 0;
    if (!h->nflagged.p) return 0;
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaDeviceSynchronize());
    int v = 0;
    CUDA_TRY(cudaMemcpy(&v, h->nflagged.p, sizeof(int), cudaMemcpyDeviceToHost));
    *h_count = v;
    return 0;
}

int b200_ivfpq_search(b200_ivfpq_t h, int64_t nq, const float* d_xq, int k, int nprobe, float* d_D, int64_t* d_I,
                      void* stream) {
    return search_impl(h, nq, d_xq, k, nprobe, nullptr, d_D, d_I, reinterpret_cast<cudaStream_t>(stream));
}

int b200_ivfpq_search_preassigned(b200_ivfpq_t h, int64_t nq, const float* d_xq, int k, int nprobe,
                                  const int64_t* d_list_ids, float* d_D, int64_t* d_I, void* stream) {
    if (!d_list_ids) return fail(B200_IVFPQ_EINVAL, "null list ids");
    return search_impl(h, nq, d_xq, k, nprobe, d_list_ids, d_D, d_I, reinterpret_cast<cudaStream_t>(stream));
}

namespace {

int search_finish(b200_ivfpq_index* h, const uint32_t* d_thr_in, float* d_D, int64_t* d_I) {
    if (!h) return fail(B200_IVFPQ_EINVAL, "null index handle");
    if (!h->pend.valid) return fail(B200_IVFPQ_ESTATE, "search_preassigned_finish without a pending _begin");
    if (!d_D || !d_I) return fail(B200_IVFPQ_EINVAL, "null result pointer");
    CUDA_TRY(cudaSetDevice(h->device));
    auto& pe = h->pend;
    pe.valid = false;
    cudaStream_t st = pe.st;
    const bool tm = h->timing && h->ev != nullptr;
    if (tm) h->filter_timed = true;
    if (tm) CUDA_TRY(cudaEventRecord(h->ev[3], st));   // the exchange sits between pair set-up and scan
    if (st_launch_rest(pe.sp, pe.qp, pe.sb, pe.nq, h->ids, d_D, d_I, pe.st_ctas, h->num_sms, st, tm ? h->ev[6] : nullptr,
                       tm ? h->ev[7] : nullptr, d_thr_in))
        return fail(B200_IVFPQ_ECUDA, "streaming scan launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    g_launches.fetch_add(3 + (d_thr_in ? 1 : 0));
    QlHostParams qp = pe.qp;
    qp.guard = st_overflow_flag(pe.sb.ctr);
    qp.qflag = static_cast<const int*>(pe.sb.qflag);
    if (ql_launch_scan(pe.sp, qp, pe.ql_ctas, st))
        return fail(B200_IVFPQ_ECUDA, "per-query-table scan launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    g_launches.fetch_add(1);
    if (tm) CUDA_TRY(cudaEventRecord(h->ev[4], st));
    size_t msmem = TopK::smem_bytes(pe.k, kMergeCap);
    int rc;
    if ((rc = set_smem(merge_query_kernel, msmem))) return rc;
    merge_query_kernel<<<(unsigned)pe.nq, kThreads, msmem, st>>>(h->out_keys.as<uint64_t>(), h->out_cnt.as<int>(), pe.sp.probe,
                                                                h->offsets.as<int64_t>(), h->ids, pe.nprobe, pe.k, 1,
                                                                h->qthr.as<uint32_t>(), d_D, d_I, qp.guard, qp.qflag);
    LAUNCH_CHECK();
    if (tm) {
        CUDA_TRY(cudaEventRecord(h->ev[5], st));
        h->timed_chunks = 1;
        h->stage_valid = true;
    }
    return 0;
}

constexpr int64_t kGraphMaxNq = 64;

uint64_t graph_epoch(const b200_ivfpq_index* h) { return (h->state_epoch << 32) ^ g_ws_epoch.load(); }

// H2D (pinned) -> search -> D2H (pinned) on h->gstream; enqueue only
int enqueue_small_search(b200_ivfpq_index* h, int64_t nq, int k, int nprobe) {
    cudaStream_t st = h->gstream;
    CUDA_TRY(cudaMemcpyAsync(h->host_xq.p, h->pin_xq, sizeof(float) * nq * h->d, cudaMemcpyHostToDevice, st));
    int rc = search_impl(h, nq, h->host_xq.as<float>(), k, nprobe, nullptr, h->host_D.as<float>(),
                         h->host_I.as<int64_t>(), st);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(h->pin_D, h->host_D.p, sizeof(float) * nq * k, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(h->pin_I, h->host_I.p, sizeof(int64_t) * nq * k, cudaMemcpyDeviceToHost, st));
    return 0;
}

int search_host_small(b200_ivfpq_index* h, int64_t nq, const float* h_xq, int k, int nprobe, float* h_D,
                      int64_t* h_I) {
    if (!h->gstream) CUDA_TRY(cudaStreamCreateWithFlags(&h->gstream, cudaStreamNonBlocking));
    const size_t xq_bytes = sizeof(float) * kGraphMaxNq * h->d;
    const size_t out_elems = (size_t)kGraphMaxNq * B200_IVFPQ_MAX_K;
    if (h->pin_xq_cap < xq_bytes) {
        if (h->pin_xq) cudaFreeHost(h->pin_xq);
        CUDA_TRY(cudaMallocHost(&h->pin_xq, xq_bytes));
        h->pin_xq_cap = xq_bytes;
        g_ws_epoch.fetch_add(1);
    }
    if (h->pin_out_cap < out_elems) {
        if (h->pin_D) cudaFreeHost(h->pin_D);
        if (h->pin_I) cudaFreeHost(h->pin_I);
        CUDA_TRY(cudaMallocHost(&h->pin_D, sizeof(float) * out_elems));
        CUDA_TRY(cudaMallocHost(&h->pin_I, sizeof(int64_t) * out_elems));
        h->pin_out_cap = out_elems;
        g_ws_epoch.fetch_add(1);
    }
    int rc;
    if ((rc = h->host_xq.ensure(xq_bytes))) return rc;
    if ((rc = h->host_D.ensure(sizeof(float) * out_elems))) return rc;
    if ((rc = h->host_I.ensure(sizeof(int64_t) * out_elems))) return rc;
    memcpy(h->pin_xq, h_xq, sizeof(float) * nq * h->d);

    auto& ge = h->graphs[std::make_tuple(nq, k, nprobe)];
    if (ge.exec && ge.epoch == graph_epoch(h)) {
        CUDA_TRY(cudaGraphLaunch(ge.exec, h->gstream));
        h->last_stream = h->gstream;
        CUDA_TRY(cudaStreamSynchronize(h->gstream));
    } else {
        if (ge.exec) {
            cudaGraphExecDestroy(ge.exec);
            ge.exec = nullptr;
        }
        if ((rc = enqueue_small_search(h, nq, k, nprobe))) return rc;
        CUDA_TRY(cudaStreamSynchronize(h->gstream));
        if (++ge.seen >= 2) {
            // workspace is sized by now: record the same sequence into a graph for the next calls
            const uint64_t epoch0 = graph_epoch(h);
            const bool timing = h->timing;
            h->timing = false;
            cudaGraph_t graph = nullptr;
            if (cudaStreamBeginCapture(h->gstream, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
                int crc = enqueue_small_search(h, nq, k, nprobe);
                cudaError_t e = cudaStreamEndCapture(h->gstream, &graph);
                if (crc == 0 && e == cudaSuccess && graph && graph_epoch(h) == epoch0) {
                    if (cudaGraphInstantiate(&ge.exec, graph, 0) == cudaSuccess) ge.epoch = epoch0;
                    else ge.exec = nullptr;
                }
                if (graph) cudaGraphDestroy(graph);
            }
            cudaGetLastError();
            h->timing = timing;
        }
    }
    memcpy(h_D, h->pin_D, sizeof(float) * nq * k);
    memcpy(h_I, h->pin_I, sizeof(int64_t) * nq * k);
    return 0;
}

}  // namespace

extern "C" int b200_ivfpq_prepare_queries(b200_ivfpq_t h, int64_t nq, const float* d_xq, void* stream) {
    if (!h) return fail(B200_IVFPQ_EINVAL, "null index handle");
    if (nq <= 0 || !d_xq) return fail(B200_IVFPQ_EINVAL, "nq <= 0 or null queries");
    h->prep.valid = false;
    if (!h->cent || !h->pq || !h->has_lists) return 0;                      // the search call reports it
    if (!(h->M == 16 || h->M == 32 || h->M == 64) || h->scan_variant == 6 || !h->st_mode) return 0;   // no per-query tables
    if (nq > (int64_t(1) << 16)) return 0;                                  // larger batches are searched in chunks
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    CUDA_TRY(cudaStreamIsCapturing(st, &cs));
    if (cs != cudaStreamCaptureStatusNone || !h->side_ok()) return 0;
    int rc;
    if ((rc = ql_prepare(h, st))) return rc;
    // only into buffers a search has already sized: the first search of a handle builds its tables itself
    if (h->ql_lut.cap < sizeof(uint16_t) * (size_t)nq * 256 * h->M || h->ql_scale.cap < sizeof(float) * nq ||
        h->ql_amin.cap < sizeof(float) * nq)
        return 0;
    CUDA_TRY(cudaEventRecord(h->ev_fork, st));
    CUDA_TRY(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
    if (ql_build_query_tables(d_xq, nq, h->pq, h->ql_mu.as<float>(), h->pq_maxnorm.as<float>(), h->d, h->M, h->dsub,
                              h->ql_lut.as<uint16_t>(), h->ql_scale.as<float>(), h->ql_amin.as<float>(), h->side))
        return fail(B200_IVFPQ_ECUDA, "query table launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    g_launches.fetch_add(1);
    CUDA_TRY(cudaEventRecord(h->ev_join, h->side));
    h->prep.valid = true;
    h->prep.xq = d_xq;
    h->prep.nq = nq;
    return 0;
}

extern "C" int b200_ivfpq_search_preassigned_begin(b200_ivfpq_t h, int64_t nq, const float* d_xq, int k, int nprobe,
                                                   const int64_t* d_list_ids, int64_t boot_lo, int64_t boot_hi,
                                                   uint32_t* d_thr_out, void* stream) {
    if (!d_list_ids || !d_thr_out) return fail(B200_IVFPQ_EINVAL, "null list ids / threshold buffer");
    if (boot_lo < 0 || boot_hi < boot_lo || boot_hi > nq) return fail(B200_IVFPQ_EINVAL, "bad bootstrap slice");
    if (nq <= 0) return fail(B200_IVFPQ_EINVAL, "nq <= 0");
    SplitArgs sa{boot_lo, boot_hi, d_thr_out};
    return search_impl(h, nq, d_xq, k, nprobe, d_list_ids, nullptr, nullptr, reinterpret_cast<cudaStream_t>(stream), &sa);
}

extern "C" int b200_ivfpq_search_preassigned_finish(b200_ivfpq_t h, const uint32_t* d_thr_in, float* d_D, int64_t* d_I) {
    return search_finish(h, d_thr_in, d_D, d_I);
}

extern "C" int b200_ivfpq_search_host(b200_ivfpq_t h, int64_t nq, const float* h_xq, int k, int nprobe, float* h_D,
                                      int64_t* h_I) {
    if (!h) return fail(B200_IVFPQ_EINVAL, "null index handle");
    if (nq < 0) return fail(B200_IVFPQ_EINVAL, "nq < 0");
    if (nq == 0) return 0;
    if (!h_xq || !h_D || !h_I) return fail(B200_IVFPQ_EINVAL, "null host pointer");
    if (k < 1 || k > B200_IVFPQ_MAX_K) return fail(B200_IVFPQ_EINVAL, "k = %d out of [1, %d]", k, B200_IVFPQ_MAX_K);
    CUDA_TRY(cudaSetDevice(h->device));
    if (h->use_graph && nq <= kGraphMaxNq && h->cent && h->pq && h->has_lists && nprobe >= 1 &&
        nprobe <= B200_IVFPQ_MAX_NPROBE)
        return search_host_small(h, nq, h_xq, k, nprobe, h_D, h_I);
    int rc;
    if ((rc = h->host_xq.ensure(sizeof(float) * nq * h->d))) return rc;
    if ((rc = h->host_D.ensure(sizeof(float) * nq * k))) return rc;
    if ((rc = h->host_I.ensure(sizeof(int64_t) * nq * k))) return rc;
    cudaStream_t st = nullptr;
    CUDA_TRY(cudaMemcpyAsync(h->host_xq.p, h_xq, sizeof(float) * nq * h->d, cudaMemcpyHostToDevice, st));
    rc = search_impl(h, nq, h->host_xq.as<float>(), k, nprobe, nullptr, h->host_D.as<float>(), h->host_I.as<int64_t>(),
                     st);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(h_D, h->host_D.p, sizeof(float) * nq * k, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(h_I, h->host_I.p, sizeof(int64_t) * nq * k, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return 0;
}

int b200_ivfpq_assign_encode(b200_ivfpq_t h, int64_t n, const float* d_x, int64_t* d_list_no, uint8_t* d_codes,
                             void* stream) {
    if (!h) return fail(B200_IVFPQ_EINVAL, "null index handle");
    if (!h->cent || !h->pq) return fail(B200_IVFPQ_ESTATE, "index is not trained (set_codebooks not called)");
    if (n < 0) return fail(B200_IVFPQ_EINVAL, "n < 0");
    if (n == 0) return 0;
    if (!d_x || !d_list_no) return fail(B200_IVFPQ_EINVAL, "null pointer");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int64_t qb = coarse_chunk(h);
    for (int64_t q0 = 0; q0 < n; q0 += qb) {
        int64_t nqc = std::min<int64_t>(qb, n - q0);
        int rc = run_coarse(h, nqc, d_x + q0 * h->d, 1, nullptr, d_list_no + q0, nullptr, st, false);
        if (rc) return rc;
    }
    if (d_codes) {
        size_t smem = sizeof(float) * (256 * (size_t)h->dsub + (size_t)h->dsub * kThreads);
        int rc = set_smem(encode_kernel, smem);
        if (rc) return rc;
        dim3 grid(static_cast<unsigned>((n + kThreads - 1) / kThreads), static_cast<unsigned>(h->M));
        encode_kernel<<<grid, kThreads, smem, st>>>(d_x, h->cent, d_list_no, h->pq, n, h->d, h->M, h->dsub, d_codes);
        LAUNCH_CHECK();
    }
    return 0;
}

int b200_ivfpq_merge_shards(int nshard, int64_t nq, int k, const float* d_Ds, const int64_t* d_Is, float* d_D,
                            int64_t* d_I, void* stream) {
    if (nshard < 1 || nq < 0) return fail(B200_IVFPQ_EINVAL, "bad nshard / nq");
    if (k < 1 || k > B200_IVFPQ_MAX_K) return fail(B200_IVFPQ_EINVAL, "k = %d out of [1, %d]", k, B200_IVFPQ_MAX_K);
    if ((int64_t)nshard * k >= (int64_t(1) << 31)) return fail(B200_IVFPQ_EINVAL, "nshard * k too large");
    if (nq == 0) return 0;
    if (!d_Ds || !d_Is || !d_D || !d_I) return fail(B200_IVFPQ_EINVAL, "null pointer");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    size_t smem = TopK::smem_bytes(k, kMergeCap);
    int rc = set_smem(merge_shards_kernel, smem);
    if (rc) return rc;
    merge_shards_kernel<<<(unsigned)nq, kThreads, smem, st>>>(d_Ds, d_Is, nshard, nq, k, d_D, d_I);
    LAUNCH_CHECK();
    return 0;
}

int b200_ivfpq_merge_shards_peer(int nshard, int64_t nq, int k, const void* const* d_bufs, int64_t d_off, int64_t i_off,
                                 float* d_D, int64_t* d_I, void* stream) {
    if (nshard < 1 || nq < 0) return fail(B200_IVFPQ_EINVAL, "bad nshard / nq");
    if (k < 1 || k > B200_IVFPQ_MAX_K) return fail(B200_IVFPQ_EINVAL, "k = %d out of [1, %d]", k, B200_IVFPQ_MAX_K);
    if ((int64_t)nshard * k >= (int64_t(1) << 31)) return fail(B200_IVFPQ_EINVAL, "nshard * k too large");
    if (nq == 0) return 0;
    if (!d_bufs || !d_D || !d_I || d_off < 0 || i_off < 0 || (d_off & 3) || (i_off & 7))
        return fail(B200_IVFPQ_EINVAL, "null pointer or misaligned offset");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    size_t smem = TopK::smem_bytes(k, kMergeCap);
    int rc = set_smem(merge_shards_peer_kernel, smem);
    if (rc) return rc;
    merge_shards_peer_kernel<<<(unsigned)nq, kThreads, smem, st>>>(
        reinterpret_cast<const unsigned char* const*>(d_bufs), d_off, i_off, nshard, nq, k, d_D, d_I);
    LAUNCH_CHECK();
    return 0;
}

int b200_ivfpq_segment_sums(int64_t k, int d, const float* d_x, const int64_t* d_order, const int64_t* d_start,
                            float* d_sums, void* stream) {
    if (k < 0 || d <= 0) return fail(B200_IVFPQ_EINVAL, "bad k / d");
    if (k == 0) return 0;
    if (!d_x || !d_order || !d_start || !d_sums) return fail(B200_IVFPQ_EINVAL, "null pointer");
    if (k >= (int64_t(1) << 31)) return fail(B200_IVFPQ_EINVAL, "k too large");
    segment_sums_kernel<<<(unsigned)k, kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_x, d_order, d_start, d, d_sums);
    LAUNCH_CHECK();
    return 0;
}

int b200_ivfpq_set_stage_timing(b200_ivfpq_t h, int enable) {
    if (!h) return fail(B200_IVFPQ_EINVAL, "null index handle");
    h->timing = enable != 0;
    return 0;
}

int b200_ivfpq_get_stage_ms(b200_ivfpq_t h, float* h_ms5) {
    if (!h || !h_ms5) return fail(B200_IVFPQ_EINVAL, "null pointer");
    if (!h->stage_valid || h->timed_chunks == 0) return fail(B200_IVFPQ_ESTATE, "no timed search recorded");
    for (int i = 0; i < 5; i++) h_ms5[i] = 0.0f;
    for (int c = 0; c < h->timed_chunks; c++) {
        auto& ev = h->evs[c];
        CUDA_TRY(cudaEventSynchronize(ev[5]));
        for (int i = 0; i < 5; i++) {
            float ms = 0.0f;
            CUDA_TRY(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
            h_ms5[i] += ms;
        }
    }
    return 0;
}

int b200_ivfpq_get_filter_ms(b200_ivfpq_t h, float* h_ms) {
    if (!h || !h_ms) return fail(B200_IVFPQ_EINVAL, "null pointer");
    if (!h->stage_valid || h->timed_chunks == 0 || !h->filter_timed)
        return fail(B200_IVFPQ_ESTATE, "no timed search through the streaming filter kernel recorded");
    *h_ms = 0.0f;
    for (int c = 0; c < h->timed_chunks; c++) {
        auto& ev = h->evs[c];
        CUDA_TRY(cudaEventSynchronize(ev[7]));
        float ms = 0.0f;
        CUDA_TRY(cudaEventElapsedTime(&ms, ev[6], ev[7]));
        *h_ms += ms;
    }
    return 0;
}

int b200_ivfpq_get_filter_stats(b200_ivfpq_t h, int64_t* h_out3, int reset) {
    if (!h || !h_out3) return fail(B200_IVFPQ_EINVAL, "null pointer");
    h_out3[0] = h_out3[1] = h_out3[2] = 0;
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaDeviceSynchronize());
    if (h->st_ctr.p) {   // streaming pipeline: counters of the LAST chunk of the last search
        unsigned char raw[kStCtrBytes];
        CUDA_TRY(cudaMemcpy(raw, h->st_ctr.p, kStCtrBytes, cudaMemcpyDeviceToHost));
        unsigned int nchunks;
        int ovf;
        unsigned long long rec, ev;
        memcpy(&nchunks, raw, 4);
        memcpy(&ovf, raw + 4, 4);
        memcpy(&rec, raw + 8, 8);
        memcpy(&ev, raw + 16, 8);
        h_out3[0] = (int64_t)rec;
        h_out3[1] = (int64_t)ev;
        if (getenv("B200_IVFPQ_STREAM_DEBUG")) {
            unsigned long long kept;
            unsigned int mx, nf;
            memcpy(&kept, raw + 24, 8);
            memcpy(&mx, raw + 32, 4);
            memcpy(&nf, raw + 36, 4);
            fprintf(stderr, "[stream] chunks %u overflow %d records %llu evals %llu kept %llu max slab %u flagged queries %u\n",
                    nchunks, ovf, rec, ev, kept, mx, nf);
        }
        h_out3[2] = ovf ? -(int64_t)ovf : (int64_t)nchunks;
    }
    if (!h->ql_counters.p) return 0;
    unsigned long long v[3];
    CUDA_TRY(cudaMemcpy(v, h->ql_counters.p, sizeof(v), cudaMemcpyDeviceToHost));
    if (!h->st_ctr.p) {
        for (int i = 0; i < 3; i++) h_out3[i] = (int64_t)v[i];
    } else if (h_out3[2] < 0) {   // the fallback launches ran: their work counts too
        for (int i = 0; i < 2; i++) h_out3[i] += (int64_t)v[i];
    }
    if (reset) CUDA_TRY(cudaMemset(h->ql_counters.p, 0, sizeof(v)));
    return 0;
}

int b200_ivfpq_get_last_scan_stats(b200_ivfpq_t h, int64_t* h_bytes, int64_t* h_codes) {
    if (!h) return fail(B200_IVFPQ_EINVAL, "null index handle");
    if (!h->stats.p) return fail(B200_IVFPQ_ESTATE, "no search recorded");
    CUDA_TRY(cudaStreamSynchronize(h->last_stream));
    PairStats s;
    CUDA_TRY(cudaMemcpy(&s, h->stats.p, sizeof(s), cudaMemcpyDeviceToHost));
    if (h_codes) *h_codes = static_cast<int64_t>(s.scan_codes);
    if (h_bytes) *h_bytes = static_cast<int64_t>(s.scan_codes) * h->M;
    return 0;
}

}  // extern "C"
