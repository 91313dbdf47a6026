// qlut.cu -- translation unit of the per-query-table filter scan (scan_qlut.cuh): kernels + their launchers.
#include <algorithm>
#include <cstdlib>

#include "qlut_api.h"
#include "scan_qlut.cuh"
#include "scan_stream.cuh"

namespace b200 {

static_assert(kQlGroupBytes == sizeof(QlGroup), "QlGroup size");
static_assert(kQlHostMaxList == kQlMaxList, "list length limit");
static_assert(kQlHostBuckets == kQlBuckets, "rank buckets");

bool ql_supported_host(int M, int d, int k) { return ql_supported(M, d, k); }

namespace {

template <int M>
int ql_grid_t(int d, int k, int64_t npairs, int num_sms) {
    const size_t smem = ql_smem_bytes<M>(d, k);
    if (smem > 227 * 1024) return 0;
    auto kernel = scan_qlut_kernel<M>;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, QlCfg<M>::kT, smem) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        return 0;
    }
    int64_t grid = static_cast<int64_t>(per_sm) * num_sms;
    if (grid > npairs) grid = npairs;
    return static_cast<int>(grid < 1 ? 1 : grid);
}

template <int M>
int ql_launch_scan_t(const ScanParams& sp, const QlParams& ql, int grid, cudaStream_t st) {
    const size_t smem = ql_smem_bytes<M>(sp.d, sp.k);
    auto kernel = scan_qlut_kernel<M>;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    kernel<<<(unsigned)grid, QlCfg<M>::kT, smem, st>>>(sp, ql);
    return cudaPeekAtLastError() == cudaSuccess ? 0 : -1;
}

template <int M>
int ql_query_tables_t(const float* xq, int64_t nq, const float* pq, const float* mu, const float* pq_maxnorm, int d,
                      int dsub, uint16_t* qlut, float* qscale, float* qamin, cudaStream_t st) {
    const char* dbg = getenv("B200_IVFPQ_QL_QMAX");   // experiments: coarser table entries (results do not change)
    const unsigned dbg_qmax = dbg ? (unsigned)atoi(dbg) : 0u;
    const uint32_t qmax = dbg_qmax ? std::min<uint32_t>(dbg_qmax, QlCfg<M>::kQMax) : QlCfg<M>::kQMax;
    constexpr int QB = 64 / M;
    const size_t smem = sizeof(float) * QB * d;
    auto kernel = ql_query_tables_kernel<M>;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return -1;
    kernel<<<(unsigned)((nq + QB - 1) / QB), 256, smem, st>>>(xq, nq, pq, mu, pq_maxnorm, d, dsub, qmax, qlut, qscale, qamin);
    return cudaPeekAtLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace

int ql_grid(int M, int d, int k, int64_t npairs, int num_sms) {
    switch (M) {
        case 16: return ql_grid_t<16>(d, k, npairs, num_sms);
        case 32: return ql_grid_t<32>(d, k, npairs, num_sms);
        case 64: return ql_grid_t<64>(d, k, npairs, num_sms);
        default: return 0;
    }
}

int ql_build_mean(const float* cent, int64_t nlist, int d, float* mu, cudaStream_t st) {
    ql_mean_kernel<<<(unsigned)d, 256, 0, st>>>(cent, nlist, d, mu);
    return cudaPeekAtLastError() == cudaSuccess ? 0 : -1;
}

int ql_build_index_data(const float* cent, const float* pq, const float* mu, const int64_t* offsets,
                        const uint8_t* codes, int64_t nlist, int d, int M, int dsub, uint16_t* snorm, float* sbmin,
                        float* sbstep, int num_sms, cudaStream_t st) {
    const size_t smem = sizeof(double) * M * 256;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(ql_sb_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return -1;
    int64_t grid = std::min<int64_t>(nlist, 8ll * num_sms);
    ql_sb_build_kernel<<<(unsigned)grid, 256, smem, st>>>(cent, pq, mu, offsets, codes, nlist, d, M, dsub, snorm, sbmin,
                                                          sbstep);
    return cudaPeekAtLastError() == cudaSuccess ? 0 : -1;
}

int ql_build_query_tables(const float* xq, int64_t nq, const float* pq, const float* mu, const float* pq_maxnorm, int d,
                          int M, int dsub, uint16_t* qlut, float* qscale, float* qamin, cudaStream_t st) {
    switch (M) {
        case 16: return ql_query_tables_t<16>(xq, nq, pq, mu, pq_maxnorm, d, dsub, qlut, qscale, qamin, st);
        case 32: return ql_query_tables_t<32>(xq, nq, pq, mu, pq_maxnorm, d, dsub, qlut, qscale, qamin, st);
        case 64: return ql_query_tables_t<64>(xq, nq, pq, mu, pq_maxnorm, d, dsub, qlut, qscale, qamin, st);
        default: return -1;
    }
}

int ql_launch_hist(const int32_t* probe, int64_t npairs, int nprobe, int64_t nlist, int nbuckets, const int64_t* offsets, int* hist,
                   PairStats* stats, cudaStream_t st) {
    ql_pair_hist_kernel<<<(unsigned)((npairs + 255) / 256), 256, 0, st>>>(probe, npairs, nprobe, nlist, nbuckets, offsets, hist, stats);
    return cudaPeekAtLastError() == cudaSuccess ? 0 : -1;
}

int ql_launch_scatter(const int32_t* probe, int64_t npairs, int nprobe, int64_t nlist, int nbuckets, const int64_t* offsets,
                      const int* start, const int* gstart, int* cursor, int32_t* order, void* groups, int gsz,
                      cudaStream_t st) {
    ql_pair_scatter_kernel<<<(unsigned)((npairs + 255) / 256), 256, 0, st>>>(probe, npairs, nprobe, nlist, nbuckets, offsets,
                                                                           start, gstart, cursor, order,
                                                                           static_cast<QlGroup*>(groups), gsz);
    return cudaPeekAtLastError() == cudaSuccess ? 0 : -1;
}

int ql_launch_scan(const ScanParams& sp, const QlHostParams& qp, int grid, cudaStream_t st) {
    QlParams ql;
    ql.snorm = qp.snorm;
    ql.sbmin = qp.sbmin;
    ql.sbstep = qp.sbstep;
    ql.pmax = qp.pmax;
    ql.qlut = qp.qlut;
    ql.qscale = qp.qscale;
    ql.qamin = qp.qamin;
    ql.counters = qp.counters;
    ql.guard = qp.guard;
    ql.qflag = qp.qflag;
    switch (sp.M) {
        case 16: return ql_launch_scan_t<16>(sp, ql, grid, st);
        case 32: return ql_launch_scan_t<32>(sp, ql, grid, st);
        case 64: return ql_launch_scan_t<64>(sp, ql, grid, st);
        default: return -1;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
static_assert(sizeof(StCounters) <= kStCtrBytes, "StCounters size");
static_assert(kStChunkRecords == kStChunk, "chunk size");

const int* st_overflow_flag(const void* ctr) { return &static_cast<const StCounters*>(ctr)->overflow; }

namespace {

template <int M>
int st_filter_grid_t(int64_t npairs, int num_sms) {
    const size_t smem = st_filter_smem<M>();
    auto kernel = st_filter_kernel<M>;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, QlCfg<M>::kT, smem) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        return 0;
    }
    int64_t grid = static_cast<int64_t>(per_sm) * num_sms;
    if (grid > npairs) grid = npairs;
    return static_cast<int>(grid < 1 ? 1 : grid);
}

template <int M>
int st_boot_t(const ScanParams& sp, const QlParams& ql, const StParams& stp, int64_t nq, int64_t lo, int64_t hi,
              cudaStream_t st) {
    if (sp.k <= 128 && getenv("B200_IVFPQ_BOOT_BLOCK") == nullptr) {
        // warp-level bootstrap: W winners per warp, 8 W >= 2 k exact candidates
        const int W = sp.k <= 32 ? 8 : sp.k <= 64 ? 16 : 32;
        const size_t bsm = st_bootw_smem<M>(sp.d, sp.nprobe, W);
#define ST_BOOTW(WW)                                                                                                 \
    {                                                                                                                \
        auto kernel = st_boot_warp_kernel<M, WW>;                                                                    \
        if (bsm > 48 * 1024 &&                                                                                       \
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsm) != cudaSuccess)      \
            return -1;                                                                                               \
        kernel<<<(unsigned)nq, kStBootThreads, bsm, st>>>(sp.xq, sp.cent, sp.pq, sp.offsets, sp.codes, sp.probe,     \
                                                          sp.nprobe, sp.d, sp.dsub, sp.k, ql, stp, sp.qthr, lo, hi); \
    }
        if (W == 8) ST_BOOTW(8) else if (W == 16) ST_BOOTW(16) else ST_BOOTW(32)
#undef ST_BOOTW
    } else {
        const size_t bsm = st_boot_smem(sp.d, M, sp.nprobe, sp.k);
        if (bsm > 48 * 1024 &&
            cudaFuncSetAttribute(st_boot_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsm) != cudaSuccess)
            return -1;
        st_boot_kernel<M><<<(unsigned)nq, kStBootThreads, bsm, st>>>(sp.xq, sp.cent, sp.pq, sp.offsets, sp.codes, sp.probe,
                                                                     sp.nprobe, sp.d, sp.dsub, sp.k, ql, stp, sp.qthr, lo, hi);
    }
    return cudaPeekAtLastError() == cudaSuccess ? 0 : -1;
}

template <int M>
int st_rest_t(const ScanParams& sp, const QlParams& ql, const StParams& stp, int64_t nq, int filter_grid, int num_sms,
              cudaStream_t st, cudaEvent_t ev0, cudaEvent_t ev1) {
    if (ev0 && cudaEventRecord(ev0, st) != cudaSuccess) return -1;
    if (M == 16 && stp.gsz == 2 && stp.two_kind != 2) {
        // lists probed by one or two queries: 32-bit table words, the four-query kernel's register pipeline
        constexpr int M16 = 16;
        const size_t fsm = st_filter_smem<M16>();
        if (cudaFuncSetAttribute(st_filter_kernel<M16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm) != cudaSuccess)
            return -1;
        st_filter_kernel<M16, true><<<(unsigned)filter_grid, QlCfg<M16>::kT, fsm, st>>>(sp, ql, stp);
    } else if (M == 16 && stp.gsz == 2) {
        // opt-in experiment: 32-bit table words, bulk-async code tiles
        const size_t fsm = st_filter2_smem();
        if (cudaFuncSetAttribute(st_filter2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm) != cudaSuccess) return -1;
        st_filter2_kernel<<<(unsigned)filter_grid, 256, fsm, st>>>(sp, ql, stp);
    } else {
        const size_t fsm = st_filter_smem<M>();
        if (cudaFuncSetAttribute(st_filter_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm) != cudaSuccess) return -1;
        st_filter_kernel<M><<<(unsigned)filter_grid, QlCfg<M>::kT, fsm, st>>>(sp, ql, stp);
    }
    if (cudaPeekAtLastError() != cudaSuccess) return -1;
    if (ev1 && cudaEventRecord(ev1, st) != cudaSuccess) return -1;
    {
        const size_t esm = st_eval_smem(M, sp.dsub);
        if (esm <= kStEvalSmemMax) {
            // the PQ codebook fits shared memory: one CTA per SM, codebook rows gathered with LDS
            if (cudaFuncSetAttribute(st_eval_kernel<M, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)esm) != cudaSuccess)
                return -1;
            st_eval_kernel<M, true><<<(unsigned)num_sms, kStEvalThreads, esm, st>>>(sp, stp);
        } else {
            st_eval_kernel<M, false><<<(unsigned)(2 * num_sms), kStEvalThreads, 0, st>>>(sp, stp);
        }
    }
    if (cudaPeekAtLastError() != cudaSuccess) return -1;
    const size_t ssm = TopK::smem_bytes(sp.k, kStSelCap);
    if (ssm > 48 * 1024 &&
        cudaFuncSetAttribute(st_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm) != cudaSuccess)
        return -1;
    st_select_kernel<<<(unsigned)nq, kStSelThreads, ssm, st>>>(stp, sp.probe, sp.offsets, sp.nprobe, sp.k);
    return cudaPeekAtLastError() == cudaSuccess ? 0 : -1;
}

void st_fill(const QlHostParams& qp, const StHostBuffers& sb, const int64_t* ids, float* D, int64_t* I, QlParams& ql,
             StParams& stp) {
    ql.snorm = qp.snorm;
    ql.sbmin = qp.sbmin;
    ql.sbstep = qp.sbstep;
    ql.pmax = qp.pmax;
    ql.qlut = qp.qlut;
    ql.qscale = qp.qscale;
    ql.qamin = qp.qamin;
    ql.counters = nullptr;
    ql.guard = nullptr;
    ql.qflag = nullptr;
    stp.srec = static_cast<uint2*>(sb.srec);
    stp.sfill = static_cast<unsigned int*>(sb.sfill);
    stp.max_chunks = sb.max_chunks;
    stp.ctr = static_cast<StCounters*>(sb.ctr);
    stp.slab = static_cast<uint64_t*>(sb.slab);
    stp.qcnt = static_cast<unsigned int*>(sb.qcnt);
    stp.qflag = static_cast<int*>(sb.qflag);
    stp.qkey = static_cast<uint64_t*>(sb.qkey);
    stp.capq = sb.capq;
    stp.gsz = sb.gsz;
    stp.two_kind = sb.two_kind;
    stp.prefix = static_cast<uint32_t*>(sb.prefix);
    stp.pdis = static_cast<float*>(sb.pdis);
    stp.ids = ids;
    {
        const char* bc = getenv("B200_IVFPQ_BOOT_CODES");
        stp.boot_codes = bc ? std::max(64, std::min(kStBootCodes, atoi(bc))) : kStBootCodes;
    }
    stp.D = D;
    stp.I = I;
}

}  // namespace

int st_filter2_grid(int64_t npairs, int num_sms) {
    const size_t smem = st_filter2_smem();
    if (cudaFuncSetAttribute(st_filter2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, st_filter2_kernel, 256, smem) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        return 0;
    }
    int64_t grid = static_cast<int64_t>(per_sm) * num_sms;
    if (grid > npairs) grid = npairs;
    return static_cast<int>(grid < 1 ? 1 : grid);
}

int st_filter_grid(int M, int64_t npairs, int num_sms) {
    switch (M) {
        case 16: return st_filter_grid_t<16>(npairs, num_sms);
        case 32: return st_filter_grid_t<32>(npairs, num_sms);
        case 64: return st_filter_grid_t<64>(npairs, num_sms);
        default: return 0;
    }
}

int st_launch_boot(const ScanParams& sp, const QlHostParams& qp, const StHostBuffers& sb, int64_t nq, int64_t lo, int64_t hi,
                   cudaStream_t st) {
    QlParams ql;
    StParams stp;
    st_fill(qp, sb, nullptr, nullptr, nullptr, ql, stp);
    if (cudaMemsetAsync(sb.ctr, 0, kStCtrBytes, st) != cudaSuccess) return -1;
    switch (sp.M) {
        case 16: return st_boot_t<16>(sp, ql, stp, nq, lo, hi, st);
        case 32: return st_boot_t<32>(sp, ql, stp, nq, lo, hi, st);
        case 64: return st_boot_t<64>(sp, ql, stp, nq, lo, hi, st);
        default: return -1;
    }
}

int st_launch_rest(const ScanParams& sp, const QlHostParams& qp, const StHostBuffers& sb, int64_t nq, const int64_t* ids,
                   float* D, int64_t* I, int filter_grid, int num_sms, cudaStream_t st, cudaEvent_t ev0, cudaEvent_t ev1,
                   const uint32_t* thr_in) {
    QlParams ql;
    StParams stp;
    st_fill(qp, sb, ids, D, I, ql, stp);
    if (thr_in) {
        st_apply_thresholds_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(thr_in, nq, sp.qthr, stp.qkey);
        if (cudaPeekAtLastError() != cudaSuccess) return -1;
    }
    switch (sp.M) {
        case 16: return st_rest_t<16>(sp, ql, stp, nq, filter_grid, num_sms, st, ev0, ev1);
        case 32: return st_rest_t<32>(sp, ql, stp, nq, filter_grid, num_sms, st, ev0, ev1);
        case 64: return st_rest_t<64>(sp, ql, stp, nq, filter_grid, num_sms, st, ev0, ev1);
        default: return -1;
    }
}

int st_launch(const ScanParams& sp, const QlHostParams& qp, const StHostBuffers& sb, int64_t nq, const int64_t* ids, float* D,
              int64_t* I, int filter_grid, int num_sms, cudaStream_t st, cudaEvent_t ev0, cudaEvent_t ev1) {
    if (st_launch_boot(sp, qp, sb, nq, 0, nq, st)) return -1;
    return st_launch_rest(sp, qp, sb, nq, ids, D, I, filter_grid, num_sms, st, ev0, ev1, nullptr);
}

}  // namespace b200
