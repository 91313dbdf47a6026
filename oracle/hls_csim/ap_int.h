// oracle/hls_csim/ap_int.h -- TEST INFRASTRUCTURE ONLY.
//
// Minimal stand-in (ours, written from scratch) for the part of Xilinx's <ap_int.h> that the reference's HLS kernels
// use, so that those kernels can be compiled with g++ and run as a C simulation (oracle/ref_fpga_shim.cpp).  Vitis HLS
// is not installed here.  Covered: ap_uint<N> for N <= 512 with zero-initialisation, construction from / conversion to
// an unsigned integer, use as an array index, and bit-range reads / writes `x.range(hi, lo)` of up to 64 bits.
// Like the real type, the object is exactly the value's bytes, little-endian (sizeof(ap_uint<32>) == 4,
// sizeof(ap_uint<512>) == 64): the kernels reinterpret `float` <-> `ap_uint<32>` through pointers and index DRAM arrays
// of ap_uint<512> / ap_uint<64> by element.
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

namespace ap_detail {
constexpr int bytes_for(int n) { return n <= 8 ? 1 : n <= 16 ? 2 : n <= 32 ? 4 : 8 * ((n + 63) / 64); }
constexpr int align_for(int n) { return bytes_for(n) < 8 ? bytes_for(n) : 8; }
}  // namespace ap_detail

template <int N>
struct alignas(ap_detail::align_for(N)) ap_uint {
    static_assert(N >= 1 && N <= 512, "ap_uint stand-in: 1..512 bits");
    static constexpr int kBytes = ap_detail::bytes_for(N);
    unsigned char b[kBytes];

    ap_uint() { std::memset(b, 0, kBytes); }
    ap_uint(unsigned long long v) {
        std::memset(b, 0, kBytes);
        if (N < 64) v &= (1ull << N) - 1ull;
        std::memcpy(b, &v, kBytes < 8 ? kBytes : 8);
    }
    ap_uint(int v) : ap_uint(static_cast<unsigned long long>(static_cast<long long>(v))) {}
    ap_uint(unsigned v) : ap_uint(static_cast<unsigned long long>(v)) {}
    ap_uint(long v) : ap_uint(static_cast<unsigned long long>(v)) {}
    ap_uint(unsigned long v) : ap_uint(static_cast<unsigned long long>(v)) {}
    template <int W>
    ap_uint(const ap_range_ref<W>& r) : ap_uint(static_cast<unsigned long long>(r)) {}

    operator unsigned long long() const {
        unsigned long long v = 0;
        std::memcpy(&v, b, kBytes < 8 ? kBytes : 8);
        return v;
    }

    unsigned long long get_range(int hi, int lo) const {
        const int width = hi - lo + 1;          // 1..64, inside the object
        unsigned long long v = 0;
        for (int i = 0; i < width; i++) {
            const int bit = lo + i;
            v |= static_cast<unsigned long long>((b[bit >> 3] >> (bit & 7)) & 1u) << i;
        }
        return v;
    }
    void set_range(int hi, int lo, unsigned long long v) {
        const int width = hi - lo + 1;
        for (int i = 0; i < width; i++) {
            const int bit = lo + i;
            const unsigned char m = static_cast<unsigned char>(1u << (bit & 7));
            if ((v >> i) & 1ull) b[bit >> 3] |= m; else b[bit >> 3] &= static_cast<unsigned char>(~m);
        }
    }
    ap_range_ref<N> range(int hi, int lo) { return ap_range_ref<N>{this, hi, lo}; }
    unsigned long long range(int hi, int lo) const { return get_range(hi, lo); }
};
