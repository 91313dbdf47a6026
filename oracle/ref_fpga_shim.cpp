// oracle/ref_fpga_shim.cpp -- TEST INFRASTRUCTURE ONLY (see ivfpq_oracle.c header).
//
// Runs the REFERENCE'S OWN HLS kernels for rows a2-a4 of the hot path as a C simulation:
//   LUT_construction.hpp : LUT_construction_wrapper  (residual q - c, T[m][k] = sum_j (r_j - pq[m][k][j])^2)
//   ADC.hpp              : PQ_lookup_computation      (dist = sum_{b<M} LUT[b][code[b]], b ascending from 0)
// and, for the selection rule of row a5,
//   priority_queue_L1.hpp: Priority_queue_L1<PQ_out_t, k, Collect_smallest>::insert_wrapper -- the systolic queue whose
//                          compare_swap (:65-75) moves an entry only past a strictly larger one.  (The accelerator's
//                          full hierarchy truncates its L1 queues and is approximate -- constants.hpp:23-31 -- which is
//                          why only the single queue, fed every candidate, is used here.)
// of retrieval_accelerator/entire_accelerator_final_<DATASET>_M<m>/src, included from where they lie under
// /root/reference (oracle/Makefile passes -I; D and M come from that directory's constants.hpp).  Nothing of the
// reference is copied.  Xilinx's <ap_int.h> / <hls_stream.h> are not installed here; oracle/hls_csim/ holds our own
// minimal stand-ins with C-simulation semantics (unbounded FIFOs, every dataflow stage run to completion).
// Built without FMA contraction: the FPGA's floating-point cores round every multiply and every add separately.
#include <cstdint>
#include <cstring>

#include "LUT_construction.hpp"
#include "ADC.hpp"
#include "priority_queue_L1.hpp"

namespace {
void push_vector_512(hls::stream<ap_uint<512> >& s, const float* v) {
    // wire format of host.cpp: a D-float vector in 512-bit words, zero-padded (size_query_vector words)
    const int words = D * 4 % 64 == 0 ? D * 4 / 64 : D * 4 / 64 + 1;
    for (int i = 0; i < words; i++) {
        float lane[16];
        for (int j = 0; j < 16; j++) lane[j] = (i * 16 + j < D) ? v[i * 16 + j] : 0.0f;
        ap_uint<512> reg;
        std::memcpy(reg.b, lane, 64);
        s.write(reg);
    }
}
}  // namespace

extern "C" __attribute__((visibility("default"))) void ref_fpga_dims(int* d, int* m) {
    *d = D;
    *m = M;
}

// pq [M][256][D/M] (Faiss order); xq [nq][D]; centers [nq][nprobe][D] (the probed cells' centroids, in probe order);
// nscan [nq][nprobe]; codes: the scanned lists' codes concatenated in (query, probe, entry) order, M bytes each.
// Out: lut [nq][nprobe][256][M] (the kernels' row format: one row per code value, M columns);
//      dist [sum nscan] in the same order as `codes`.
extern "C" __attribute__((visibility("default")))
int ref_fpga_lut_adc(int nq, int nprobe, const float* pq, const float* xq, const float* centers, const int* nscan,
                     const uint8_t* codes, float* lut, float* dist) {
    hls::stream<float> s_pq;
    hls::stream<ap_uint<512> > s_q, s_c;
    hls::stream<distance_LUT_parallel_t> s_lut, s_lut_fwd;
    hls::stream<PQ_in_t> s_codes;
    hls::stream<int> s_nscan;
    hls::stream<PQ_out_t> s_res;

    for (long i = 0; i < static_cast<long>(M) * LUT_ENTRY_NUM * (D / M); i++) s_pq.write(pq[i]);
    long total = 0;
    for (int q = 0; q < nq; q++) {
        push_vector_512(s_q, xq + static_cast<long>(q) * D);
        for (int p = 0; p < nprobe; p++) {
            push_vector_512(s_c, centers + (static_cast<long>(q) * nprobe + p) * D);
            const int n = nscan[q * nprobe + p];
            s_nscan.write(n);
            for (int e = 0; e < n; e++, total++) {
                PQ_in_t in;
                in.valid = true;
                in.cell_ID = p;
                in.offset = e;
                for (int b = 0; b < M; b++) in.PQ_code[b] = ap_uint<8>(static_cast<unsigned>(codes[total * M + b]));
                s_codes.write(in);
            }
        }
    }
    LUT_construction_wrapper(nq, nprobe, s_pq, s_q, s_c, s_lut);
    PQ_lookup_computation(nq, nprobe, s_lut, s_codes, s_nscan, s_lut_fwd, s_res);
    for (long r = 0; r < static_cast<long>(nq) * nprobe * LUT_ENTRY_NUM; r++) {
        const distance_LUT_parallel_t row = s_lut_fwd.read();
        std::memcpy(lut + r * M, row.dist, sizeof(float) * M);
    }
    for (long i = 0; i < total; i++) dist[i] = s_res.read().dist;
    return (s_lut.empty() && s_codes.empty() && s_res.empty()) ? 0 : -1;
}

namespace {
template <int QS>
void run_queue(int n, const float* dist, int* out_off, float* out_dist) {
    hls::stream<int> s_iter;
    hls::stream<PQ_out_t> s_in, s_out;
    s_iter.write(n);
    for (int i = 0; i < n; i++) {
        PQ_out_t e;
        e.cell_ID = 0;
        e.offset = i;          // scan order
        e.dist = dist[i];
        s_in.write(e);
    }
    Priority_queue_L1<PQ_out_t, QS, Collect_smallest> queue;
    queue.insert_wrapper(1, s_iter, s_in, s_out);
    for (int i = 0; i < QS; i++) {          // queue order, not sorted; unfilled slots keep dist = LARGE_NUM
        const PQ_out_t e = s_out.read();
        out_off[i] = e.offset;
        out_dist[i] = e.dist;
    }
}
}  // namespace

// The reference's queue of length k (1, 10 or 100) fed n candidates in scan order; returns its k slots.
extern "C" __attribute__((visibility("default")))
int ref_fpga_queue_l1(int k, int n, const float* dist, int* out_off, float* out_dist) {
    switch (k) {
        case 1: run_queue<1>(n, dist, out_off, out_dist); return 0;
        case 10: run_queue<10>(n, dist, out_off, out_dist); return 0;
        case 100: run_queue<100>(n, dist, out_off, out_dist); return 0;
        default: return -1;
    }
}
