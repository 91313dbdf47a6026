// select_radix.cuh -- exact k-selection over one row of non-negative floats by MSB radix select.
//
// Used for the coarse quantizer's per-query select (K1b: nprobe, or the L pre-filter candidates, smallest of nlist
// distances).  Keys are the 64-bit (distance bits << 32) | index, all distinct, ordered like the oracle's
// (distance, centroid id): "sort, take nprobe", ties -> lower id (IVFPQ_1B_search.ipynb:7997-7999).
//
// One CTA per row.  Up to 8 passes of 8 bits find the k-th smallest key exactly: each pass histograms the next byte
// of the keys that still match the prefix (warp-aggregated shared-memory atomics), one warp scans the 256 bins for
// the bucket where the cumulative count crosses k.  After the four distance bytes, if every key that equals the
// pivot distance is needed, the four index passes are skipped (the common case: no tie at the boundary).  Then all
// keys <= pivot (exactly k of them) are collected and bitonic-sorted for the ordered output.  Cost is ~5 passes
// over the row instead of sorting thousands of candidates.
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

#include "topk.cuh"

namespace b200 {

__host__ __device__ inline int select_pow2(int k) {
    int p = 1;
    while (p < k) p <<= 1;
    return p;
}

__host__ __device__ inline size_t select_smem_bytes(int k) {
    return sizeof(uint64_t) * static_cast<size_t>(select_pow2(k)) + sizeof(int) * (256 + 8);
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
select_radix_kernel(const float* __restrict__ dist, int64_t n, int64_t stride, int k, int32_t* __restrict__ ids32,
                    int64_t* __restrict__ ids64, float* __restrict__ dis_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int P = select_pow2(k);
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);       // [P]
    int* hist = reinterpret_cast<int*>(keys + P);                 // [256]
    int* ctrl = hist + 256;   // [0] remaining  [1] out count  [2] prefix hi  [3] prefix lo  [4] done flag
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t q = blockIdx.x;
    const float* row = dist + q * stride;
    const int kk = static_cast<int>(n < k ? n : k);               // number of real results

    if (tid == 0) {
        ctrl[0] = kk;
        ctrl[1] = 0;
        ctrl[2] = 0;
        ctrl[3] = 0;
        ctrl[4] = 0;
    }
    uint64_t pivot = ~0ull;                                        // n <= k: everything is selected
    if (n > k) {
        uint64_t prefix = 0;                                       // bytes above the current one, right-aligned
        for (int b = 7; b >= 0; b--) {
            for (int i = tid; i < 256; i += THREADS) hist[i] = 0;
            __syncthreads();
            for (int64_t base = 0; base < n; base += THREADS) {
                const int64_t i = base + tid;
                int digit = 256;                                   // sentinel: not counted
                if (i < n) {
                    const uint64_t key = make_key(__float_as_uint(row[i]), static_cast<uint32_t>(i));
                    const bool match = (b == 7) || ((key >> (8 * (b + 1))) == prefix);
                    if (match) digit = static_cast<int>((key >> (8 * b)) & 255u);
                }
                const unsigned peers = __match_any_sync(0xffffffffu, digit);
                if (digit < 256 && lane == __ffs(peers) - 1) atomicAdd(&hist[digit], __popc(peers));
            }
            __syncthreads();
            if (warp == 0) {
                // bins 8*lane .. 8*lane+7; find the first bin whose cumulative count reaches `remaining`
                const int remaining = ctrl[0];
                int local[8], sum = 0;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    local[j] = hist[lane * 8 + j];
                    sum += local[j];
                }
                int incl = sum;
                for (int o = 1; o < 32; o <<= 1) {
                    int y = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += y;
                }
                const int excl = incl - sum;
                const unsigned hit = __ballot_sync(0xffffffffu, incl >= remaining);
                const int owner = __ffs(hit) - 1;
                if (lane == owner) {
                    int cum = excl, bin = 0;
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        if (cum + local[j] >= remaining) {
                            bin = j;
                            break;
                        }
                        cum += local[j];
                    }
                    const int digit = lane * 8 + bin;
                    ctrl[0] = remaining - cum;                     // still needed inside the chosen bucket
                    const uint64_t np = (prefix << 8) | static_cast<uint64_t>(digit);
                    ctrl[2] = static_cast<int>(np >> 32);
                    ctrl[3] = static_cast<int>(np & 0xffffffffu);
                    // after the 4 distance bytes: if the whole bucket (all keys with the pivot distance) is
                    // needed, the index bytes do not matter
                    ctrl[4] = (b == 4 && local[bin] == remaining - cum) ? 1 : 0;
                }
            }
            __syncthreads();
            prefix = (static_cast<uint64_t>(static_cast<uint32_t>(ctrl[2])) << 32) | static_cast<uint32_t>(ctrl[3]);
            if (b == 4 && ctrl[4]) {
                pivot = (prefix << 32) | 0xffffffffull;
                break;
            }
            if (b == 0) pivot = prefix;
        }
    }
    __syncthreads();
    // collect every key <= pivot: exactly kk of them
    for (int i = tid; i < P; i += THREADS) keys[i] = kPadKey;
    __syncthreads();
    for (int64_t base = 0; base < n; base += THREADS) {
        const int64_t i = base + tid;
        bool take = false;
        uint64_t key = 0;
        if (i < n) {
            key = make_key(__float_as_uint(row[i]), static_cast<uint32_t>(i));
            take = key <= pivot;
        }
        const unsigned mask = __ballot_sync(0xffffffffu, take);
        if (mask) {
            const int leader = __ffs(mask) - 1;
            int slot = 0;
            if (lane == leader) slot = atomicAdd(&ctrl[1], __popc(mask));
            slot = __shfl_sync(0xffffffffu, slot, leader);
            if (take) {
                slot += __popc(mask & lanemask_lt());
                if (slot < P) keys[slot] = key;                    // NaN rows can exceed; guarded
            }
        }
    }
    __syncthreads();
    for (int size = 2; size <= P; size <<= 1) {
        for (int strd = size >> 1; strd > 0; strd >>= 1) {
            for (int t = tid; t < (P >> 1); t += THREADS) {
                const int lo = 2 * t - (t & (strd - 1));
                const int hi = lo + strd;
                const bool asc = (lo & size) == 0;
                const uint64_t a = keys[lo], b2 = keys[hi];
                if ((a > b2) == asc) {
                    keys[lo] = b2;
                    keys[hi] = a;
                }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < k; i += THREADS) {
        int32_t id = -1;
        float dv = FLT_MAX;
        const uint64_t key = i < P ? keys[i] : kPadKey;
        // +inf / NaN distances are never results (the oracle's FLT_MAX / -1 convention for unfilled slots)
        if (i < kk && key != kPadKey && static_cast<uint32_t>(key >> 32) < kInfBits) {
            id = static_cast<int32_t>(key & 0xffffffffu);
            dv = __uint_as_float(static_cast<uint32_t>(key >> 32));
        }
        if (ids32) ids32[q * k + i] = id;
        if (ids64) ids64[q * k + i] = id;
        if (dis_out) dis_out[q * k + i] = dv;
    }
}

}  // namespace b200
