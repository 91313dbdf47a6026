// topk.cuh -- exact block-level k-selection over 64-bit keys, shared by every selection on the path:
//   coarse select   key = (distance bits << 32) | centroid id          (ties -> lower id)
//   scan top-k      key = (distance bits << 32) | offset in list       (ties -> earlier offset)
//   per-query merge key = (distance bits << 32) | (probe rank * k + j) (ties -> earlier probe)
//   shard merge     key = (distance bits << 32) | (shard * k + j)
// Distances are sums of squares, so their fp32 bit patterns order like unsigned integers and the
// 64-bit unsigned compare is exactly the oracle's (distance, tag) total order
// (oracle/ivfpq_oracle.c ent_less; reference compare rule priority_queue_L1.hpp:65-75, strict <).
//
// Mechanism: candidates that pass the running threshold (distance <= current k-th best) are appended
// to a shared-memory queue with one warp-aggregated atomic; the queue is folded into the sorted best
// list only when it could overflow (bitonic sort of the queue + merge-by-rank), so the common case per
// candidate is one compare.  Keys must be distinct (they are: the tag is unique).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b200 {

constexpr uint32_t kInfBits = 0x7f800000u;   // +inf: "no threshold yet"
constexpr uint64_t kPadKey = ~0ull;

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

__device__ __forceinline__ uint64_t make_key(uint32_t dist_bits, uint32_t tag) {
    return (static_cast<uint64_t>(dist_bits) << 32) | tag;
}

struct TopK {
    uint64_t* best;   // [2 * k] ping-pong sorted lists
    uint64_t* queue;  // [cap]
    int* meta;        // [0] nbest  [1] qcount  [2] cur  [3] threshold bits
    int k;
    int cap;

    static __host__ __device__ size_t smem_bytes(int k, int cap) {
        return sizeof(uint64_t) * (2 * static_cast<size_t>(k) + cap) + 4 * sizeof(int);
    }

    // carve out of a shared-memory region (8-byte aligned)
    __device__ void bind(void* smem, int k_, int cap_) {
        k = k_;
        cap = cap_;
        best = reinterpret_cast<uint64_t*>(smem);
        queue = best + 2 * k_;
        meta = reinterpret_cast<int*>(queue + cap_);
    }

    // one thread; caller syncs afterwards
    __device__ void reset(uint32_t thr_bits) {
        meta[0] = 0;
        meta[1] = 0;
        meta[2] = 0;
        meta[3] = static_cast<int>(thr_bits);
    }

    __device__ __forceinline__ uint32_t threshold() const { return static_cast<uint32_t>(meta[3]); }
    __device__ __forceinline__ int pending() const { return meta[1]; }
    __device__ __forceinline__ int count() const { return meta[0]; }
    __device__ __forceinline__ const uint64_t* sorted() const { return best + meta[2] * k; }

    // whole warp calls (converged); lanes with pred append their key.  Capacity is the caller's
    // invariant: flush() is called whenever pending() + (candidates of the next tile) > cap.
    __device__ __forceinline__ void push(bool pred, uint64_t key) {
        unsigned mask = __ballot_sync(0xffffffffu, pred);
        if (mask == 0) return;
        int lane = threadIdx.x & 31;
        int leader = __ffs(mask) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(&meta[1], __popc(mask));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (pred) queue[base + __popc(mask & lanemask_lt())] = key;
    }

    // Barrier + uniform flush decision in one step.  Every thread evaluates "pending > limit" after its own
    // pushes and the barrier ORs the votes: the last thread to vote sees every push, so all threads get the
    // same answer (a plain read after a __syncthreads() would race with fast threads already pushing the
    // next tile).  Returns after the queue has room for another tile.
    template <int THREADS>
    __device__ __forceinline__ void sync_and_flush_if_over(int limit, uint32_t ext_thr) {
        const int seen = *reinterpret_cast<volatile int*>(&meta[1]);
        if (__syncthreads_or(seen > limit)) flush<THREADS>(ext_thr);
    }

    static __device__ __forceinline__ int lower_bound(const uint64_t* a, int n, uint64_t key) {
        int lo = 0, hi = n;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (a[mid] < key) lo = mid + 1;
            else hi = mid;
        }
        return lo;
    }

    // ---- warp-level sorting networks (registers + shuffles only) -----------------------------------------------
    // A run of 32 R keys lives in one warp: element e = 32 r + lane is v[r] of that lane.  Compare-exchanges with
    // stride >= 32 are register-to-register, smaller strides are butterfly shuffles.
    static __device__ __forceinline__ void cmpx(uint64_t& lo, uint64_t& hi, bool asc) {
        const bool sw = (lo > hi) == asc;
        const uint64_t a = sw ? hi : lo, b = sw ? lo : hi;
        lo = a;
        hi = b;
    }
    template <int R>
    static __device__ __forceinline__ void warp_stage(uint64_t (&v)[R], int lane, int size, int stride) {
        if (stride >= 32) {
            const int rs = stride >> 5;
#pragma unroll
            for (int r = 0; r < R; r++)
                if ((r & rs) == 0) cmpx(v[r], v[r + rs], ((r << 5) & size) == 0);
        } else {
#pragma unroll
            for (int r = 0; r < R; r++) {
                const uint64_t other = __shfl_xor_sync(0xffffffffu, v[r], stride);
                const bool asc = ((((r << 5) | lane) & size) == 0);
                const bool keep_min = ((lane & stride) == 0) == asc;
                v[r] = ((v[r] < other) == keep_min) ? v[r] : other;
            }
        }
    }
    template <int R>
    static __device__ __forceinline__ void warp_sort(uint64_t (&v)[R], int lane) {
#pragma unroll
        for (int size = 2; size <= 32 * R; size <<= 1)
#pragma unroll
            for (int stride = size >> 1; stride > 0; stride >>= 1) warp_stage<R>(v, lane, size, stride);
    }
    // a, b ascending runs -> a = the 32 R smallest of their union, ascending
    template <int R>
    static __device__ __forceinline__ void warp_lower(uint64_t (&a)[R], const uint64_t (&b)[R], int lane) {
#pragma unroll
        for (int r = 0; r < R; r++) {
            const uint64_t brev = __shfl_sync(0xffffffffu, b[R - 1 - r], 31 - lane);
            a[r] = a[r] < brev ? a[r] : brev;
        }
#pragma unroll
        for (int stride = 16 * R; stride > 0; stride >>= 1) warp_stage<R>(a, lane, 64 * R, stride);   // all ascending
    }
    static __device__ __forceinline__ uint64_t warp_sort32(uint64_t key, int lane) {
        uint64_t v[1] = {key};
        warp_sort<1>(v, lane);
        return v[0];
    }

    // n > 32 R queue entries -> queue[0 .. 32 R) = their 32 R smallest, ascending.  Every warp reduces its share of
    // the queue to one sorted run in registers ("sort a chunk, merge, keep the lower half"), warp 0 merges the runs:
    // three barriers instead of one per stage of a shared-memory bitonic sort of the whole queue.
    // (a free-standing function, not inlined: the sorting networks are large and run rarely compared with the scan
    // loops they would otherwise be inlined into several times)
    template <int R, int THREADS>
    static __device__ __noinline__ void reduce_queue(uint64_t* __restrict__ queue, int n) {
        constexpr int NW = THREADS / 32, RUN = 32 * R;
        const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
        uint64_t acc[R];
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = kPadKey;
        for (int base = w * RUN; base < n; base += NW * RUN) {
            uint64_t v[R];
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int i = base + 32 * r + lane;
                v[r] = i < n ? queue[i] : kPadKey;
            }
            warp_sort<R>(v, lane);
            warp_lower<R>(acc, v, lane);
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < R; r++) queue[w * RUN + 32 * r + lane] = acc[r];
        __syncthreads();
        if (w == 0) {
            const int nruns = min(NW, (n + RUN - 1) / RUN);   // warps beyond that hold only padding
            for (int j = 1; j < nruns; j++) {
                uint64_t v[R];
#pragma unroll
                for (int r = 0; r < R; r++) v[r] = queue[j * RUN + 32 * r + lane];
                warp_lower<R>(acc, v, lane);
            }
#pragma unroll
            for (int r = 0; r < R; r++) queue[32 * r + lane] = acc[r];
        }
        __syncthreads();
    }

    // all THREADS threads call, after a __syncthreads() that made the queue writes visible.
    // ext_thr: an externally known upper bound on the k-th best distance (kInfBits if none).
    template <int THREADS>
    __device__ void flush(uint32_t ext_thr) {
        const int tid = threadIdx.x;
        const int n = meta[1];
        if (n == 0) return;   // uniform
        if (n <= 32) {
            // the common case once a threshold is known: one warp sorts the queue in registers (shuffles only)
            if (tid < 32) {
                const uint64_t key = warp_sort32(tid < n ? queue[tid] : kPadKey, tid);
                queue[tid] = key;
            }
            __syncthreads();
        } else if (k <= 32 && THREADS <= cap) {
            reduce_queue<1, THREADS>(queue, n);
        } else if (k <= 64 && 2 * THREADS <= cap) {
            reduce_queue<2, THREADS>(queue, n);
        } else if (k <= 128 && 4 * THREADS <= cap) {
            reduce_queue<4, THREADS>(queue, n);
        } else {
            int P = 1;
            while (P < n) P <<= 1;
            for (int i = n + tid; i < P; i += THREADS) queue[i] = kPadKey;
            __syncthreads();
            for (int size = 2; size <= P; size <<= 1) {
                for (int stride = size >> 1; stride > 0; stride >>= 1) {
                    for (int t = tid; t < (P >> 1); t += THREADS) {
                        int lo = 2 * t - (t & (stride - 1));
                        int hi = lo + stride;
                        bool asc = (lo & size) == 0;
                        uint64_t a = queue[lo], b = queue[hi];
                        if ((a > b) == asc) {
                            queue[lo] = b;
                            queue[hi] = a;
                        }
                    }
                    __syncthreads();
                }
            }
        }
        const int cur = meta[2], nb = meta[0];
        const uint64_t* old = best + cur * k;
        uint64_t* out = best + (cur ^ 1) * k;
        const int nnew = n < k ? n : k;   // queue is sorted: only its first k can survive
        for (int i = tid; i < nnew; i += THREADS) {
            uint64_t key = queue[i];
            int r = i + lower_bound(old, nb, key);
            if (r < k) out[r] = key;
        }
        for (int t = tid; t < nb; t += THREADS) {
            uint64_t key = old[t];
            int r = t + lower_bound(queue, nnew, key);
            if (r < k) out[r] = key;
        }
        __syncthreads();
        if (tid == 0) {
            int nn = nb + n < k ? nb + n : k;
            meta[0] = nn;
            meta[1] = 0;
            meta[2] = cur ^ 1;
            uint32_t thr = ext_thr;
            if (nn == k) {
                uint32_t kth = static_cast<uint32_t>(out[k - 1] >> 32);
                thr = kth < thr ? kth : thr;
            }
            meta[3] = static_cast<int>(thr);
        }
        __syncthreads();
    }
};

// Fold NQ queues at once when every one of them holds at most 32 pending keys (the common case once thresholds are
// known): warp q sorts queue q in registers, THREADS / NQ threads merge it into its best list -- three barriers for
// all queues together instead of three per queue.  Returns false, having done nothing, when some queue holds more
// (the caller then folds them one by one).  All THREADS threads call, after a barrier; THREADS >= 32 NQ.
template <int THREADS, int NQ>
__device__ __forceinline__ bool topk_fold_small(TopK (&tk)[NQ], const uint32_t (&ext)[NQ]) {
    static_assert(THREADS % NQ == 0 && THREADS / 32 >= NQ, "one warp per queue");
    constexpr int PER = THREADS / NQ;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, part = tid / PER, tl = tid % PER;
    int n[NQ];
    bool small = true, any = false;
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        n[q] = tk[q].meta[1];
        small = small && n[q] <= 32;
        any = any || n[q] > 0;
    }
    if (!small) return false;   // uniform: every thread read the same counters
    if (!any) return true;
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        if (w == q && n[q] > 0) {
            const uint64_t key = TopK::warp_sort32(lane < n[q] ? tk[q].queue[lane] : kPadKey, lane);
            tk[q].queue[lane] = key;
        }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        if (part == q && n[q] > 0) {
            const int k = tk[q].k, cur = tk[q].meta[2], nb = tk[q].meta[0];
            const uint64_t* old = tk[q].best + cur * k;
            uint64_t* out = tk[q].best + (cur ^ 1) * k;
            const uint64_t* queue = tk[q].queue;
            const int nnew = n[q] < k ? n[q] : k;
            for (int i = tl; i < nnew; i += PER) {
                const uint64_t key = queue[i];
                const int r = i + TopK::lower_bound(old, nb, key);
                if (r < k) out[r] = key;
            }
            for (int t = tl; t < nb; t += PER) {
                const uint64_t key = old[t];
                const int r = t + TopK::lower_bound(queue, nnew, key);
                if (r < k) out[r] = key;
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        if (tid == q * PER && n[q] > 0) {
            const int k = tk[q].k, cur = tk[q].meta[2], nb = tk[q].meta[0];
            const uint64_t* out = tk[q].best + (cur ^ 1) * k;
            const int nn = nb + n[q] < k ? nb + n[q] : k;
            tk[q].meta[0] = nn;
            tk[q].meta[1] = 0;
            tk[q].meta[2] = cur ^ 1;
            uint32_t thr = ext[q];
            if (nn == k) {
                const uint32_t kth = static_cast<uint32_t>(out[k - 1] >> 32);
                thr = kth < thr ? kth : thr;
            }
            tk[q].meta[3] = static_cast<int>(thr);
        }
    }
    __syncthreads();
    return true;
}

}  // namespace b200
