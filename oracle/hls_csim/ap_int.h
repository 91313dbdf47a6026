// oracle/hls_csim/ap_int.h -- TEST INFRASTRUCTURE ONLY.
//
// Minimal stand-in (ours, written from scratch) for the part of Xilinx's <ap_int.h> that the reference's HLS kernels
// use, so that those kernels can be compiled with g++ and run as a C simulation (oracle/ref_fpga_shim.cpp).  Vitis HLS
// is not installed here.  Covered: ap_uint<N> for N <= 512 with zero-initialisation, construction from / conversion to
// an unsigned integer, use as an array index, and bit-range reads / writes `x.range(hi, lo)` of up to 64 bits.  The
// value lives little-endian in 64-bit words at offset 0 of the object, which is what the kernels' own
// `*((float*) &ap_uint32)` reinterpretation relies on (as it does with the real header).
#pragma once
#include <cstdint>
#include <cstring>

template <int N>
struct ap_uint;

template <int N>
struct ap_range_ref {
    ap_uint<N>* obj;
    int hi, lo;
    operator unsigned long long() const { return obj->get_range(hi, lo); }
    ap_range_ref& operator=(unsigned long long v) {
        obj->set_range(hi, lo, v);
        return *this;
    }
    template <int W>
    ap_range_ref& operator=(const ap_uint<W>& v) {
        obj->set_range(hi, lo, static_cast<unsigned long long>(v));
        return *this;
    }
    ap_range_ref& operator=(const ap_range_ref& o) {
        obj->set_range(hi, lo, static_cast<unsigned long long>(o));
        return *this;
    }
};

template <int N>
struct ap_uint {
    static_assert(N >= 1 && N <= 512, "ap_uint stand-in: 1..512 bits");
    static constexpr int kWords = (N + 63) / 64;
    uint64_t w[kWords];

    ap_uint() { std::memset(w, 0, sizeof(w)); }
    ap_uint(unsigned long long v) {
        std::memset(w, 0, sizeof(w));
        w[0] = N >= 64 ? v : (v & ((1ull << (N % 64)) - 1ull));
    }
    ap_uint(int v) : ap_uint(static_cast<unsigned long long>(static_cast<long long>(v))) {}
    ap_uint(unsigned v) : ap_uint(static_cast<unsigned long long>(v)) {}
    ap_uint(long v) : ap_uint(static_cast<unsigned long long>(v)) {}
    ap_uint(unsigned long v) : ap_uint(static_cast<unsigned long long>(v)) {}
    template <int W>
    ap_uint(const ap_range_ref<W>& r) : ap_uint(static_cast<unsigned long long>(r)) {}

    operator unsigned long long() const { return w[0]; }

    unsigned long long get_range(int hi, int lo) const {
        const int width = hi - lo + 1;          // 1..64, inside the object
        unsigned long long v = 0;
        for (int b = 0; b < width; b++) {
            const int bit = lo + b;
            v |= ((w[bit >> 6] >> (bit & 63)) & 1ull) << b;
        }
        return v;
    }
    void set_range(int hi, int lo, unsigned long long v) {
        const int width = hi - lo + 1;
        for (int b = 0; b < width; b++) {
            const int bit = lo + b;
            const uint64_t m = 1ull << (bit & 63);
            if ((v >> b) & 1ull) w[bit >> 6] |= m; else w[bit >> 6] &= ~m;
        }
    }
    ap_range_ref<N> range(int hi, int lo) { return ap_range_ref<N>{this, hi, lo}; }
    unsigned long long range(int hi, int lo) const { return get_range(hi, lo); }
};
