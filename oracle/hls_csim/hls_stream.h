// oracle/hls_csim/hls_stream.h -- TEST INFRASTRUCTURE ONLY.
//
// Minimal stand-in (ours) for Xilinx's <hls_stream.h> with C-simulation semantics: an unbounded FIFO.  A dataflow
// region of HLS functions is then run one function after the other, each to completion (what Vitis csim does).
// Reading an empty stream is a hard error here (it would be a deadlock in hardware).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <deque>

namespace hls {
template <typename T>
class stream {
public:
    stream() {}
    explicit stream(const char*) {}
    void write(const T& v) { q_.push_back(v); }
    T read() {
        if (q_.empty()) {
            std::fprintf(stderr, "hls_csim: read from an empty stream\n");
            std::abort();
        }
        T v = q_.front();
        q_.pop_front();
        return v;
    }
    bool empty() const { return q_.empty(); }
    size_t size() const { return q_.size(); }

private:
    std::deque<T> q_;
};
}  // namespace hls
