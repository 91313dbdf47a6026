// qlut_api.h -- host entry points of the per-query-table filter scan (scan_qlut.cuh), compiled in qlut.cu.
// All pointers are device pointers unless named h_*; every function returns 0, or -1 after a CUDA error (the caller
// reads cudaGetLastError / cudaPeekAtLastError).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "scan_types.cuh"

namespace b200 {

struct QlHostParams {
    // per index (built by ql_build_index_data)
    const uint16_t* snorm;
    const float* sbmin;
    const float* sbstep;
    float pmax;
    // per batch (built by ql_build_query_tables)
    const uint16_t* qlut;
    const float* qscale;
    const float* qamin;
    unsigned long long* counters;   // 3 counters or nullptr
    const int* guard;               // fallback launch: the kernel runs only if *guard != 0 (nullptr: always)
    const int* qflag;               // fallback launch with (*guard & 3) == 0: only the queries with qflag[q] != 0
};

constexpr uint32_t kQlHostMaxList = 1u << 28;
constexpr size_t kQlGroupBytes = 48;

bool ql_supported_host(int M, int d, int k);
// CTAs to launch (resident CTAs per SM x SMs, capped by the number of pairs); 0 when the kernel does not fit
int ql_grid(int M, int d, int k, int64_t npairs, int num_sms);
// mu (d floats) = mean of the coarse centroids
int ql_build_mean(const float* cent, int64_t nlist, int d, float* mu, cudaStream_t st);
// snorm (ntotal u16), sbmin / sbstep (nlist f32) from the codes
int ql_build_index_data(const float* cent, const float* pq, const float* mu, const int64_t* offsets,
                        const uint8_t* codes, int64_t nlist, int d, int M, int dsub, uint16_t* snorm, float* sbmin,
                        float* sbstep, int num_sms, cudaStream_t st);
// qlut (nq, 256, M) u16, qscale / qamin (nq) f32
int ql_build_query_tables(const float* xq, int64_t nq, const float* pq, const float* mu, const float* pq_maxnorm, int d,
                          int M, int dsub, uint16_t* qlut, float* qscale, float* qamin, cudaStream_t st);
// pair setup with the key (rank bucket, list): hist / start / gstart / cursor hold kQlHostBuckets * nlist entries
constexpr int kQlHostBuckets = 3;
int ql_launch_hist(const int32_t* probe, int64_t npairs, int nprobe, int64_t nlist, int nbuckets, const int64_t* offsets, int* hist,
                   PairStats* stats, cudaStream_t st);
int ql_launch_scatter(const int32_t* probe, int64_t npairs, int nprobe, int64_t nlist, int nbuckets, const int64_t* offsets,
                      const int* start, const int* gstart, int* cursor, int32_t* order, void* groups, int gsz,
                      cudaStream_t st);
int ql_launch_scan(const ScanParams& sp, const QlHostParams& qp, int grid, cudaStream_t st);

// ---- streaming pipeline (scan_stream.cuh): bootstrap thresholds -> filter -> exact evaluation -> per-query select ----
struct StHostBuffers {
    void* srec;               // max_chunks * 64 * 8 bytes
    void* sfill;              // max_chunks * 4 bytes
    unsigned int max_chunks;
    void* ctr;                // kStCtrBytes, zeroed by st_launch
    void* slab;               // nq * capq * 8 bytes
    void* qcnt;               // nq * 4 bytes
    void* qflag;              // nq * 4 bytes
    void* qkey;               // nq * 8 bytes
    int capq;
    int gsz;                  // pairs per work item the groups were built with (4, or 2 for the two-query filters)
    int two_kind;             // gsz == 2: 1 = st_filter_kernel<16, true> (LDS.32 tables), 2 = st_filter2_kernel (bulk-async)
    void* prefix;             // nq * nprobe * 4 bytes
    void* pdis;               // nq * nprobe * 4 bytes
};
constexpr size_t kStCtrBytes = 48;
constexpr int kStChunkRecords = 64;
int st_filter_grid(int M, int64_t npairs, int num_sms);        // 0: does not fit
int st_filter2_grid(int64_t npairs, int num_sms);             // the two-query filter (M = 16)
// enqueues the four kernels; the overflow flag (an int, != 0 after an overflow) lives at st_overflow_flag(ctr)
// ev0 / ev1 (optional): recorded right before / after the filter kernel
int st_launch(const ScanParams& sp, const QlHostParams& qp, const StHostBuffers& sb, int64_t nq, const int64_t* ids, float* D,
              int64_t* I, int filter_grid, int num_sms, cudaStream_t st, cudaEvent_t ev0 = nullptr, cudaEvent_t ev1 = nullptr);
const int* st_overflow_flag(const void* ctr);
// the same pipeline in two halves, for the multi-GPU threshold exchange: boot bootstraps the thresholds of the queries
// [lo, hi) only (scan positions and coarse distances of all of them); rest takes thr_in (nq distance bits, or nullptr)
// -- e.g. the all-reduce MIN over the ranks of every rank's qthr -- before filter / evaluation / select
int st_launch_boot(const ScanParams& sp, const QlHostParams& qp, const StHostBuffers& sb, int64_t nq, int64_t lo, int64_t hi,
                   cudaStream_t st);
int st_launch_rest(const ScanParams& sp, const QlHostParams& qp, const StHostBuffers& sb, int64_t nq, const int64_t* ids,
                   float* D, int64_t* I, int filter_grid, int num_sms, cudaStream_t st, cudaEvent_t ev0, cudaEvent_t ev1,
                   const uint32_t* thr_in);

}  // namespace b200
