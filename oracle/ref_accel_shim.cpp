// oracle/ref_accel_shim.cpp -- TEST INFRASTRUCTURE ONLY (see ivfpq_oracle.c header).
//
// Runs the REFERENCE'S COMPLETE ACCELERATOR KERNEL as a C simulation: the top level `vadd`
// (retrieval_accelerator/entire_accelerator_final_<DATASET>_M<m>/src/vadd.cpp) with everything it instantiates --
// network input parsing, LUT construction, the ADC PEs over four DRAM banks, the hierarchical priority queue with the
// vector-id lookup, result packing -- included from where it lies under /root/reference (oracle/Makefile passes -I).
// Nothing of the reference is copied; the DRAM images it expects (meta data, query packets, PQ-code banks, vector-id
// banks) are laid out by OUR host code in oracle/ivfpq_oracle.py from the formats the kernel parses.  Xilinx's
// ap_int.h / hls_stream.h are our stand-ins (oracle/hls_csim/); the dataflow region runs stage by stage.
//
// The accelerator is an APPROXIMATE top-k by design: each of its 2 x ADC_PE_NUM first-level queues keeps only
// PRIORITY_QUEUE_LEN_L1 entries (constants.hpp:23-31).  Callers use it on inputs where no first-level queue overflows
// with winners, so that its answer is the exact top-TOPK.
#include <cstdint>

#include "vadd.cpp"

extern "C" __attribute__((visibility("default")))
void ref_accel_dims(int* d, int* m, int* topk, int* adc_pe_num, int* l1_len) {
    *d = D;
    *m = M;
    *topk = TOPK;
    *adc_pe_num = ADC_PE_NUM;
    *l1_len = PRIORITY_QUEUE_LEN_L1;
}

extern "C" __attribute__((visibility("default")))
void ref_accel_run(int query_num, int nlist, int nprobe, int* meta_data_init, void* in_dram, const void* pq0,
                   const void* pq1, const void* pq2, const void* pq3, void* vid0, void* vid1, void* vid2, void* vid3,
                   void* out_dram) {
    static_assert(sizeof(ap_uint<512>) == 64 && sizeof(ap_uint<64>) == 8, "DRAM word sizes");
    vadd(query_num, nlist, nprobe, meta_data_init, static_cast<ap_uint<512>*>(in_dram),
         static_cast<const ap_uint<512>*>(pq0), static_cast<const ap_uint<512>*>(pq1),
         static_cast<const ap_uint<512>*>(pq2), static_cast<const ap_uint<512>*>(pq3),
         static_cast<ap_uint<64>*>(vid0), static_cast<ap_uint<64>*>(vid1), static_cast<ap_uint<64>*>(vid2),
         static_cast<ap_uint<64>*>(vid3), static_cast<ap_uint<512>*>(out_dram));
}
