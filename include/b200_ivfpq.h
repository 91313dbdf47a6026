/*
 * b200_ivfpq.h -- C-ABI of the B200-native IVF-PQ search engine (libb200ivfpq.so).
 *
 * This is the drop-in boundary for the retrieval hot path of Chameleon: every entry point below is
 * what a binding of the reference's Faiss calls for this path would bind.  The reference reaches the
 * path through Faiss's SWIG layer (third party, not vendored); the call sites it replaces are cited
 * per function as file:line under /root/reference/Chameleon/.
 *
 * Conventions
 *   - plain C, no torch / C++ types; all sizes explicit.
 *   - pointers named d_* are DEVICE pointers (e.g. a torch tensor's data_ptr()); h_* are HOST pointers.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  Device-pointer entry
 *     points are asynchronous on that stream; *_host entry points synchronise before returning.
 *   - every function returns 0 on success or a B200_IVFPQ_E* code; b200_ivfpq_last_error() returns the
 *     message of the last failure on the calling thread (Faiss raises RuntimeError from FAISS_THROW;
 *     the Python layer turns a non-zero code into RuntimeError with this message).
 *   - arithmetic contract: BASELINE.md section 2 / oracle/ivfpq_oracle.c header (fp32, non-fused,
 *     sequential; coarse ties -> lower centroid id; top-k total order (distance, probe rank, list
 *     offset); unfilled slots id -1 / distance FLT_MAX).
 *   - layouts: centroids (nlist, d) f32 row-major; pq (M, 256, dsub) f32
 *     (Faiss_experiments/my_faiss_extract_scripts/extract_Enzian_U250_required_data.py:222-246);
 *     inverted lists flattened CSR: list l = rows [offsets[l], offsets[l+1]) of codes (ntotal, M) u8 and
 *     ids (ntotal) i64, insertion order inside a list (same file :264-279, Faiss ArrayInvertedLists).
 */
#ifndef B200_IVFPQ_H
#define B200_IVFPQ_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define B200_API __attribute__((visibility("default")))
#else
#define B200_API
#endif

#define B200_IVFPQ_OK 0
#define B200_IVFPQ_EINVAL 1     /* bad argument */
#define B200_IVFPQ_ESTATE 2     /* index not trained / lists not set */
#define B200_IVFPQ_ECUDA 3      /* CUDA runtime error (message carries cudaGetErrorString) */
#define B200_IVFPQ_ENOMEM 4     /* workspace allocation failed */
#define B200_IVFPQ_EUNSUPPORTED 5

#define B200_IVFPQ_MAX_K 2048      /* same bound Faiss-GPU applies to k and nprobe */
#define B200_IVFPQ_MAX_NPROBE 2048

typedef struct b200_ivfpq_index* b200_ivfpq_t;

/* library / device ------------------------------------------------------------------------------ */
B200_API const char* b200_ivfpq_last_error(void);
B200_API const char* b200_ivfpq_version(void);
/* number of kernel launches this library has issued on the calling process (bench.py gpu_launches) */
B200_API int64_t b200_ivfpq_launch_count(void);

/* lifecycle -- faiss.IndexIVFPQ(quantizer, d, nlist, m, nbits) / index_factory(d, "IVF<nlist>,PQ<m>")
 * (Faiss_experiments/IVFPQ_random_dataset.py:24, bench_cpu_performance.py:98).  nbits must be 8. */
B200_API int b200_ivfpq_create(int d, int64_t nlist, int m, int nbits, b200_ivfpq_t* out);
B200_API int b200_ivfpq_destroy(b200_ivfpq_t h);

/* trained state: coarse centroids + PQ codebook.  The pointers are BORROWED (caller keeps them alive);
 * this is what index.quantizer / index.pq.centroids hold after index.train(x)
 * (bench_cpu_performance.py:100-109). */
B200_API int b200_ivfpq_set_codebooks(b200_ivfpq_t h, const float* d_centroids, const float* d_pq);

/* populated state: CSR inverted lists, BORROWED.  h_offsets is a HOST array of nlist+1 entries and is
 * copied; d_ids may be NULL (id = position).  Mirrors invlists.list_size/get_codes/get_ids
 * (extract_Enzian_U250_required_data.py:264-279). */
B200_API int b200_ivfpq_set_lists(b200_ivfpq_t h, const int64_t* h_offsets, const uint8_t* d_codes, const int64_t* d_ids,
                         int64_t ntotal);

/* a1 -- quantizer.search(xq, nprobe) (llm_inference_gpu/ralm/index_scanner/index_scanner.py:73,
 * ralm/retriever/faiss_retriever.py:259).  Outputs (nq, nprobe): ids i64, distances f32, ascending. */
B200_API int b200_ivfpq_coarse(b200_ivfpq_t h, int64_t nq, const float* d_xq, int nprobe, int64_t* d_ids, float* d_dis,
                      void* stream);

/* K1 internals, exposed for tests and profiling: the tensor-core pre-filter scores s(q, c) = ||c||^2 - 2 q.c
 * (tcgen05 split-bf16 GEMM) into d_scores (nq, nlist) f32, and the number of queries of all b200_ivfpq_coarse /
 * search calls so far whose candidate set could not be proven sufficient and were redone by the exact kernels
 * (synchronises the device).  coarse_scores returns B200_IVFPQ_EUNSUPPORTED when the tensor-core path is off. */
B200_API int b200_ivfpq_coarse_scores(b200_ivfpq_t h, int64_t nq, const float* d_xq, float* d_scores, void* stream);
B200_API int b200_ivfpq_coarse_fallbacks(b200_ivfpq_t h, int64_t* h_count);

/* a1..a6 -- index.search(xq, k) with index.nprobe = nprobe (bench_cpu_performance.py:269,
 * faiss_retriever.py:254).  Outputs (nq, k): D f32 ascending, I i64. */
B200_API int b200_ivfpq_search(b200_ivfpq_t h, int64_t nq, const float* d_xq, int k, int nprobe, float* d_D, int64_t* d_I,
                      void* stream);

/* a2..a6 -- faiss.contrib.ivf_tools.search_preassigned(index, xq, k, list_ids)
 * (ralm/server/faiss_server.py:233, faiss_retriever.py:265).  d_list_ids is (nq, nprobe) i64, entries < 0
 * are skipped. */
B200_API int b200_ivfpq_search_preassigned(b200_ivfpq_t h, int64_t nq, const float* d_xq, int k, int nprobe,
                                  const int64_t* d_list_ids, float* d_D, int64_t* d_I, void* stream);

/* Optional head start for the NEXT search of exactly (nq, d_xq) on this handle: the per-query tables the scan filters
 * with depend on the queries only, so they can be built (on the handle's own side stream, ordered after `stream`) while
 * the caller still computes or exchanges the probe lists -- the sharded search of bench_gpu_performance_OSDI.py:586-604
 * runs the coarse quantizer on a slice of the batch and all-gathers the probes first.  The queries must not change
 * between this call and the search.  A no-op (return 0) wherever it does not apply; results never depend on it. */
B200_API int b200_ivfpq_prepare_queries(b200_ivfpq_t h, int64_t nq, const float* d_xq, void* stream);

/* The same search in two halves, for an index sharded by vector over several GPUs (bench_gpu_performance_OSDI.py:586-604,
 * co.shard = True).  Every shard needs, per query, an upper bound on the FINAL k-th distance before it filters its codes;
 * any shard's own k-th best distance is one.  _begin does everything up to those bootstrap thresholds, but only for the
 * queries [boot_lo, boot_hi) (d_thr_out, nq x u32 distance bits, +inf bits elsewhere); the caller combines the ranks'
 * arrays (an all-reduce MIN of nq x 4 bytes over NVLink) and hands the result to _finish, which filters, evaluates and
 * selects on the stream given to _begin.  One pending search per handle; B200_IVFPQ_EUNSUPPORTED when the streaming
 * pipeline does not apply to this shape (use b200_ivfpq_search_preassigned then). */
B200_API int b200_ivfpq_search_preassigned_begin(b200_ivfpq_t h, int64_t nq, const float* d_xq, int k, int nprobe,
                                        const int64_t* d_list_ids, int64_t boot_lo, int64_t boot_hi,
                                        uint32_t* d_thr_out, void* stream);
B200_API int b200_ivfpq_search_preassigned_finish(b200_ivfpq_t h, const uint32_t* d_thr_in, float* d_D, int64_t* d_I);

/* Same call with HOST buffers (numpy arrays, the reference's convention: faiss_retriever.py:227-275):
 * H2D of the queries, search, D2H of the results, synchronised.  This is the end-to-end path bench.py
 * reports as e2e. */
B200_API int b200_ivfpq_search_host(b200_ivfpq_t h, int64_t nq, const float* h_xq, int k, int nprobe, float* h_D,
                           int64_t* h_I);

/* a9 -- the device half of index.add / add_with_ids (bench_cpu_performance.py:159): nearest centroid
 * (ties -> lower id) and PQ code of the residual (ties -> lower code).  Outputs: d_list_no (n) i64,
 * d_codes (n, M) u8.  Appending to the lists is host logic (Python layer). */
B200_API int b200_ivfpq_assign_encode(b200_ivfpq_t h, int64_t n, const float* d_x, int64_t* d_list_no, uint8_t* d_codes,
                             void* stream);

/* a10 -- the centroid update of index.train's Lloyd iterations (bench_cpu_performance.py:98-109, bench_gpu_1bn.py:
 * 522-542): d_sums[c] = sum of the rows d_x[d_order[i]] for i in [d_start[c], d_start[c+1]), added one after the other
 * in that order (fp32), so that training is reproducible bit for bit.  d_x is (n, d) f32, d_order i64, d_start (k+1)
 * i64, d_sums (k, d) f32. */
B200_API int b200_ivfpq_segment_sums(int64_t k, int d, const float* d_x, const int64_t* d_order, const int64_t* d_start,
                            float* d_sums, void* stream);

/* multi-GPU merge (K5) -- what Faiss IndexShards does on the host after per-GPU searches
 * (bench_gpu_performance_OSDI.py:587-604; merge semantics bench_multi_cpu_performance_OSDI.py:203-219):
 * d_Ds / d_Is are (nshard, nq, k) as produced by an all-gather of per-shard results; output (nq, k) is the
 * k smallest under (distance, shard, position).  Index-free: usable on any rank. */
B200_API int b200_ivfpq_merge_shards(int nshard, int64_t nq, int k, const float* d_Ds, const int64_t* d_Is, float* d_D,
                            int64_t* d_I, void* stream);

/* The same merge over PEER MEMORY: d_bufs is a DEVICE array of nshard base pointers, entry s being rank s's result
 * buffer mapped into this GPU's address space (CUDA peer access over NVLink / NVSwitch, e.g. torch symmetric memory's
 * buffer_ptrs_dev); every buffer holds D (nq, k) f32 at byte offset d_off and I (nq, k) i64 at i_off.  The kernel's
 * loads are the exchange: no all-gather, no staging copy.  The caller orders it against the producers (a
 * symmetric-memory barrier before, and one after before the buffers are overwritten). */
B200_API int b200_ivfpq_merge_shards_peer(int nshard, int64_t nq, int k, const void* const* d_bufs, int64_t d_off,
                                 int64_t i_off, float* d_D, int64_t* d_I, void* stream);

/* instrumentation for bench.py / ncu: device time (ms, CUDA events on `stream`) the last search spent in
 * each stage: [0] coarse distances, [1] coarse select, [2] pair setup, [3] LUT+scan+top-k, [4] merge.
 * Timing is off by default; enabling it adds event records but no synchronisation to the search call
 * (the getter synchronises on the events). */
B200_API int b200_ivfpq_set_stage_timing(b200_ivfpq_t h, int enable);
B200_API int b200_ivfpq_get_stage_ms(b200_ivfpq_t h, float* h_ms5);
/* algorithmic scan bytes (sum over probed lists of list_size * M) and codes scanned by the last search;
 * synchronises the stream of that search. */
B200_API int b200_ivfpq_get_last_scan_stats(b200_ivfpq_t h, int64_t* h_bytes, int64_t* h_codes);

/* device time (ms) of the streaming pipeline's filter kernel (csrc/scan_stream.cuh st_filter_kernel, the kernel the
 * roofline is quoted on) in the last timed search; B200_IVFPQ_ESTATE if that search did not run it. */
B200_API int b200_ivfpq_get_filter_ms(b200_ivfpq_t h, float* h_ms);

/* instrumentation of the per-query-table filter scan (csrc/scan_qlut.cuh, csrc/scan_stream.cuh): [0] survivor records the integer
 * filter passed, [1] exact fp32 evaluations, [2] record chunks used by the streaming pipeline in the last search, or
 * minus the overflow bits when its fallback launches answered (part of) the batch; in-kernel path (B200_IVFPQ_STREAM=0,
 * handle created with B200_IVFPQ_QL_STATS=1): work items, accumulated since the last reset.  Synchronises. */
B200_API int b200_ivfpq_get_filter_stats(b200_ivfpq_t h, int64_t* h_out3, int reset);

#ifdef __cplusplus
}
#endif
#endif /* B200_IVFPQ_H */
