/*
 * oracle/ivfpq_oracle.c -- CPU restatement of the IVF-PQ search path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker, not the product.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product path
 * (chameleon-rag-acceleration_b200/) never links, imports or calls anything in oracle/.
 *
 * What it restates.  The arithmetic of the path lives in Faiss (faiss-cpu==1.7.1, pinned at
 * Chameleon/Faiss_experiments/README.md:18), which is NOT vendored under /root/reference and not
 * installable here.  The reference does hold two independent restatements of the same algorithm,
 * and this file follows them line by line:
 *   [NB]   Chameleon/Faiss_experiments/my_faiss_extract_scripts/IVFPQ_1B_search.ipynb:7922-8028
 *          (distance_full_vec, construct_distance_table, estimate_distance(s), search_single_query),
 *          checked equal to Faiss at :8094-8095.
 *   [FPGA] Chameleon/retrieval_accelerator/entire_accelerator_final_SIFT_M32/src/
 *          LUT_construction.hpp:180-209, ADC.hpp:75-99, priority_queue_L1.hpp:65-75.
 * Parity status: the LUT arithmetic is pinned by the reference's literal known-answer test
 * (LUT_construction_PE_D128_M32/src/host.cpp:44-109; tests/golden/lut_kat_d128_m32.npz).  The
 * coarse stage is pinned against the reference's own cell-selection code run here (vendored hnswlib
 * brute force, host.cpp:516-581, via oracle/ref_coarse_shim.cpp -> oracle/_ref/; tests/test_reference_coarse.py),
 * and the LUT / ADC arithmetic against the reference's HLS kernels (LUT_construction.hpp, ADC.hpp) run as a C
 * simulation (oracle/ref_fpga_shim.cpp; bit for bit, tests/test_reference_fpga_kernels.py), and the whole search against
 * the reference's host + complete accelerator kernel (oracle/ref_accel_shim.cpp; tests/test_reference_system.py).
 * The end-to-end search result is "parity unpinned" against the Faiss binary: no runnable Faiss, no
 * SIFT1B index.
 *
 * Arithmetic contract, fixed so that CPU and GPU agree bit for bit:
 *   - everything is IEEE fp32; every multiply and every add is rounded separately (no FMA
 *     contraction: build with -ffp-contract=off);
 *   - L2^2(a,b) = sum_j (a_j-b_j)^2, accumulated sequentially for j = 0..d-1 starting from 0;
 *   - coarse: the nprobe smallest (distance, centroid id) pairs, ascending; ties -> lower id;
 *   - LUT: T[m][k] = sum_{j<dsub} ((q-c)[m*dsub+j] - pq[m][k][j])^2, same sequential form;
 *   - ADC: dist = sum_{m=0..M-1} T[m][code[m]], ascending m, starting from 0;
 *   - top-k: max-heap semantics, replace iff new < top (strict), scan order = probe rank then
 *     list offset; expressed here as the total order (dist, scan sequence number); rows ascending;
 *     unfilled slots: id -1, distance FLT_MAX (what Faiss 1.7.1 emits).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* [NB] distance_full_vec, ipynb:7922-7927.  Sequential, non-fused. */
static inline float l2sqr_seq(const float* a, const float* b, int d) {
    float acc = 0.0f;
    for (int j = 0; j < d; j++) {
        float diff = a[j] - b[j];
        float sq = diff * diff;
        acc = acc + sq;
    }
    return acc;
}

ORACLE_API float oracle_l2sqr(const float* a, const float* b, int d) { return l2sqr_seq(a, b, d); }

ORACLE_API int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

ORACLE_API void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---- (distance, tag) max-heap used for both coarse select and top-k ------------------------- */
typedef struct {
    float dis;
    int64_t tag; /* centroid id (coarse) or scan sequence number (top-k) */
    int64_t payload;
} ent_t;

static inline int ent_less(const ent_t* a, const ent_t* b) {
    return a->dis < b->dis || (a->dis == b->dis && a->tag < b->tag);
}

static void heap_sift_down(ent_t* h, int n, int i) {
    for (;;) {
        int l = 2 * i + 1, r = l + 1, big = i;
        if (l < n && ent_less(&h[big], &h[l])) big = l;
        if (r < n && ent_less(&h[big], &h[r])) big = r;
        if (big == i) return;
        ent_t t = h[i];
        h[i] = h[big];
        h[big] = t;
        i = big;
    }
}

static void heap_sift_up(ent_t* h, int i) {
    while (i > 0) {
        int p = (i - 1) / 2;
        if (!ent_less(&h[p], &h[i])) return;
        ent_t t = h[i];
        h[i] = h[p];
        h[p] = t;
        i = p;
    }
}

/* keep the k smallest under (dis, tag).  With tag = monotonically increasing scan number this is
 * exactly "replace iff dis < top" ([FPGA] priority_queue_L1.hpp:65-75, strict <). */
static inline void heap_offer(ent_t* h, int* n, int k, ent_t e) {
    if (*n < k) {
        h[*n] = e;
        heap_sift_up(h, *n);
        (*n)++;
    } else if (ent_less(&e, &h[0])) {
        h[0] = e;
        heap_sift_down(h, k, 0);
    }
}

static int ent_cmp(const void* a, const void* b) {
    const ent_t* x = (const ent_t*)a;
    const ent_t* y = (const ent_t*)b;
    if (ent_less(x, y)) return -1;
    if (ent_less(y, x)) return 1;
    return 0;
}

/* ---- a1: coarse quantizer ------------------------------------------------------------------- */
/* [NB] search_single_query, ipynb:7991-7999: L2^2 to every centroid, sort, take nprobe. */
ORACLE_API void oracle_coarse(int64_t nq, int d, const float* xq, int64_t nlist, const float* centroids,
                              int nprobe, int64_t* out_ids, float* out_dis) {
#pragma omp parallel
    {
        ent_t* heap = (ent_t*)malloc(sizeof(ent_t) * (size_t)(nprobe > 0 ? nprobe : 1));
#pragma omp for schedule(dynamic, 8)
        for (int64_t q = 0; q < nq; q++) {
            int n = 0;
            for (int64_t c = 0; c < nlist; c++) {
                ent_t e;
                e.dis = l2sqr_seq(xq + q * d, centroids + c * d, d);
                e.tag = c;
                e.payload = c;
                heap_offer(heap, &n, nprobe, e);
            }
            qsort(heap, (size_t)n, sizeof(ent_t), ent_cmp);
            for (int i = 0; i < nprobe; i++) {
                out_ids[q * nprobe + i] = i < n ? heap[i].payload : -1;
                out_dis[q * nprobe + i] = i < n ? heap[i].dis : FLT_MAX;
            }
        }
        free(heap);
    }
}

/* ---- a2 + a3: residual and LUT --------------------------------------------------------------- */
/* [NB] q_res = q_vec - coarse_cen[cell_id] (ipynb:8006), construct_distance_table (ipynb:7929-7946);
 * [FPGA] LUT_construction.hpp:182-209.  pq is (M, 256, dsub) row-major
 * (extract_Enzian_U250_required_data.py:222-232); T is (M, 256). */
ORACLE_API void oracle_lut(int d, int M, const float* q, const float* c, const float* pq, float* T) {
    int dsub = d / M;
    for (int m = 0; m < M; m++) {
        for (int k = 0; k < 256; k++) {
            const float* p = pq + ((size_t)m * 256 + k) * dsub;
            float acc = 0.0f;
            for (int j = 0; j < dsub; j++) {
                float r = q[m * dsub + j] - c[m * dsub + j];
                float diff = r - p[j];
                float sq = diff * diff;
                acc = acc + sq;
            }
            T[m * 256 + k] = acc;
        }
    }
}

/* ---- a4: ADC scan ---------------------------------------------------------------------------- */
/* [NB] estimate_distance, ipynb:7948-7960; [FPGA] ADC.hpp:88-91.  codes is (n, M) uint8. */
ORACLE_API void oracle_adc(int M, const float* T, int64_t n, const uint8_t* codes, float* dist) {
    for (int64_t i = 0; i < n; i++) {
        const uint8_t* code = codes + i * M;
        float acc = 0.0f;
        for (int m = 0; m < M; m++) acc = acc + T[m * 256 + code[m]];
        dist[i] = acc;
    }
}

/* ---- a4 + a5 + a6: scan the probed lists of every query, top-k, id lookup ------------------- */
/* [NB] search_single_query, ipynb:8001-8017.  Inverted lists are CSR: list l holds codes
 * [offsets[l], offsets[l+1]) of `codes` (list-major, M bytes each) and the matching `ids`
 * -- the flattened ArrayInvertedLists layout of extract_Enzian_U250_required_data.py:264-279.
 * probe_ids is (nq, nprobe); negative entries are skipped (Faiss search_preassigned). */
ORACLE_API void oracle_search_preassigned(int64_t nq, int d, const float* xq, int64_t nlist,
                                          const float* centroids, int M, const float* pq,
                                          const int64_t* offsets, const uint8_t* codes, const int64_t* ids,
                                          int nprobe, const int64_t* probe_ids, int k, float* D, int64_t* I) {
    (void)nlist;
#pragma omp parallel
    {
        float* T = (float*)malloc(sizeof(float) * (size_t)M * 256);
        ent_t* heap = (ent_t*)malloc(sizeof(ent_t) * (size_t)(k > 0 ? k : 1));
#pragma omp for schedule(dynamic, 4)
        for (int64_t q = 0; q < nq; q++) {
            int n = 0;
            int64_t seq = 0;
            for (int p = 0; p < nprobe; p++) {
                int64_t l = probe_ids[q * nprobe + p];
                if (l < 0) continue;
                int64_t beg = offsets[l], end = offsets[l + 1];
                if (end > beg) {
                    oracle_lut(d, M, xq + q * d, centroids + l * d, pq, T);
                    for (int64_t i = beg; i < end; i++) {
                        const uint8_t* code = codes + i * M;
                        float acc = 0.0f;
                        for (int m = 0; m < M; m++) acc = acc + T[m * 256 + code[m]];
                        if (n < k || acc < heap[0].dis) { /* strict <: earlier-scanned wins ties */
                            ent_t e;
                            e.dis = acc;
                            e.tag = seq + (i - beg);
                            e.payload = ids[i];
                            heap_offer(heap, &n, k, e);
                        }
                    }
                }
                seq += end - beg;
            }
            qsort(heap, (size_t)n, sizeof(ent_t), ent_cmp);
            for (int i = 0; i < k; i++) {
                D[q * k + i] = i < n ? heap[i].dis : FLT_MAX;
                I[q * k + i] = i < n ? heap[i].payload : -1;
            }
        }
        free(T);
        free(heap);
    }
}

/* index.search(xq, k): coarse then scan ([NB] search_batch_query, ipynb:8019-8028). */
ORACLE_API void oracle_search(int64_t nq, int d, const float* xq, int64_t nlist, const float* centroids, int M,
                              const float* pq, const int64_t* offsets, const uint8_t* codes, const int64_t* ids,
                              int nprobe, int k, float* D, int64_t* I, int64_t* probe_ids_out,
                              float* probe_dis_out) {
    int64_t* pid = probe_ids_out ? probe_ids_out : (int64_t*)malloc(sizeof(int64_t) * (size_t)(nq * nprobe));
    float* pdis = probe_dis_out ? probe_dis_out : (float*)malloc(sizeof(float) * (size_t)(nq * nprobe));
    oracle_coarse(nq, d, xq, nlist, centroids, nprobe, pid, pdis);
    oracle_search_preassigned(nq, d, xq, nlist, centroids, M, pq, offsets, codes, ids, nprobe, pid, k, D, I);
    if (!probe_ids_out) free(pid);
    if (!probe_dis_out) free(pdis);
}

/* ---- a9: encode (index.add): nearest centroid, residual, nearest sub-centroid ---------------- */
/* call sites bench_cpu_performance.py:159, run_RALM_SYN_dataset.py:278.  Ties -> lower index. */
ORACLE_API void oracle_assign(int64_t n, int d, const float* x, int64_t nlist, const float* centroids,
                              int64_t* list_no) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; i++) {
        float best = 0.0f;
        int64_t arg = -1;
        for (int64_t c = 0; c < nlist; c++) {
            float dis = l2sqr_seq(x + i * d, centroids + c * d, d);
            if (arg < 0 || dis < best) {
                best = dis;
                arg = c;
            }
        }
        list_no[i] = arg;
    }
}

ORACLE_API void oracle_encode(int64_t n, int d, const float* x, const float* centroids, const int64_t* list_no,
                              int M, const float* pq, uint8_t* codes) {
    int dsub = d / M;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; i++) {
        const float* c = centroids + list_no[i] * d;
        for (int m = 0; m < M; m++) {
            float best = 0.0f;
            int arg = -1;
            for (int k = 0; k < 256; k++) {
                const float* p = pq + ((size_t)m * 256 + k) * dsub;
                float acc = 0.0f;
                for (int j = 0; j < dsub; j++) {
                    float r = x[i * d + m * dsub + j] - c[m * dsub + j];
                    float diff = r - p[j];
                    float sq = diff * diff;
                    acc = acc + sq;
                }
                if (arg < 0 || acc < best) {
                    best = acc;
                    arg = k;
                }
            }
            codes[i * M + m] = (uint8_t)arg;
        }
    }
}

/* ---- multi-shard merge ----------------------------------------------------------------------- */
/* bench_multi_cpu_performance_OSDI.py:203-219: concatenate the per-shard (D, I) rows, stable
 * argsort by distance, take k.  Stable = (distance, shard, position).  Empty slots (id -1) lose. */
ORACLE_API void oracle_merge_shards(int nshard, int64_t nq, int k, const float* Ds, const int64_t* Is, float* D,
                                    int64_t* I) {
#pragma omp parallel
    {
        ent_t* all = (ent_t*)malloc(sizeof(ent_t) * (size_t)nshard * (size_t)k);
#pragma omp for
        for (int64_t q = 0; q < nq; q++) {
            int n = 0;
            for (int s = 0; s < nshard; s++)
                for (int i = 0; i < k; i++) {
                    int64_t id = Is[((int64_t)s * nq + q) * k + i];
                    if (id < 0) continue;
                    all[n].dis = Ds[((int64_t)s * nq + q) * k + i];
                    all[n].tag = (int64_t)s * k + i;
                    all[n].payload = id;
                    n++;
                }
            qsort(all, (size_t)n, sizeof(ent_t), ent_cmp);
            for (int i = 0; i < k; i++) {
                D[q * k + i] = i < n ? all[i].dis : FLT_MAX;
                I[q * k + i] = i < n ? all[i].payload : -1;
            }
        }
        free(all);
    }
}
