// Single-pass device-wide exclusive scan (decoupled look-back) as a kernel template.
//
// One launch scans n items: every CTA takes a tile ticket, scans its 2048 items locally, publishes the tile
// aggregate, looks back over its predecessors' status words (a whole warp reads 32 of them at a time) until it meets
// an inclusive prefix, publishes its own inclusive prefix and writes its outputs.  Replaces the three launches
// (tile scan, scan of the tile sums, add) + the one-thread "total" kernel of the first version of primitives.cu.
//
// The scanned value is a PAIR of non-negative int32 counters packed as (a << 31) | b: two scans for the price of
// one (components: cluster rank and CSR offset; means: chunk base and multi-chunk list position).  Sums of either
// counter must stay below 2^31.
//
// Op interface (a trivially copyable struct passed by value):
//   __device__ int64_t size() const;                                  item count (may read device memory)
//   __device__ unsigned long long load(int64_t i) const;              lb_pair(a, b) of item i
//   __device__ void store(int64_t i, unsigned long long excl, unsigned long long item) const;
//   __device__ void finish(unsigned long long total) const;           called once by the last tile (n == 0: tile 0)
//
// state: 1 + ceil(max items / 2048) 64-bit words, ZEROED before the launch (word 0 is the ticket counter).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace slic {

constexpr int LB_THREADS = 256;
constexpr int LB_ITEMS = 8;
constexpr int LB_TILE = LB_THREADS * LB_ITEMS;

__host__ __device__ __forceinline__ unsigned long long lb_pair(int a, int b) {
    return ((unsigned long long)(unsigned)a << 31) | (unsigned long long)(unsigned)b;
}
__host__ __device__ __forceinline__ int lb_a(unsigned long long p) { return (int)((p >> 31) & 0x7fffffffull); }
__host__ __device__ __forceinline__ int lb_b(unsigned long long p) { return (int)(p & 0x7fffffffull); }

inline size_t lookback_state_bytes(int64_t max_items) {
    const int64_t tiles = max_items > 0 ? (max_items + LB_TILE - 1) / LB_TILE : 1;
    return (size_t)(tiles + 1) * sizeof(unsigned long long);
}
inline unsigned lookback_grid(int64_t max_items) {
    return (unsigned)(max_items > 0 ? (max_items + LB_TILE - 1) / LB_TILE : 1);
}

constexpr unsigned long long LB_FLAG_AGG = 1ull << 62, LB_FLAG_PREFIX = 2ull << 62, LB_VALUE_MASK = (1ull << 62) - 1;

template <class Op>
__global__ void __launch_bounds__(LB_THREADS) lookback_scan_kernel(const Op op, unsigned long long* __restrict__ state) {
    __shared__ unsigned long long s_warp[LB_THREADS / 32];
    __shared__ unsigned long long s_excl, s_total;
    __shared__ int s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = (int)atomicAdd(state, 1ull);   // tiles start in ticket order: predecessors are running or done
    __syncthreads();
    const int tile = s_tile;
    const int64_t n = op.size();
    const int64_t num_tiles = n > 0 ? (n + LB_TILE - 1) / LB_TILE : 1;
    if (tile >= num_tiles) return;
    volatile unsigned long long* status = state + 1;

    const int64_t base = (int64_t)tile * LB_TILE + (int64_t)threadIdx.x * LB_ITEMS;
    unsigned long long v[LB_ITEMS], sum = 0;
#pragma unroll
    for (int i = 0; i < LB_ITEMS; ++i) {
        v[i] = (base + i < n) ? op.load(base + i) : 0ull;
        sum += v[i];
    }
    // block-wide exclusive scan of the per-thread sums
    unsigned long long incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const unsigned long long w = lane < LB_THREADS / 32 ? s_warp[lane] : 0ull;
        unsigned long long wi = w;
#pragma unroll
        for (int o = 1; o < LB_THREADS / 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        if (lane < LB_THREADS / 32) s_warp[lane] = wi - w;
        const unsigned long long total = __shfl_sync(0xffffffffu, wi, LB_THREADS / 32 - 1);
        // decoupled look-back, 32 predecessors per step
        unsigned long long excl = 0;
        if (tile > 0) {
            if (lane == 0) status[tile] = LB_FLAG_AGG | total;
            int t = tile - 1;
            while (true) {
                const int idx = t - lane;
                unsigned long long wv = LB_FLAG_PREFIX;   // before tile 0: an empty inclusive prefix
                if (idx >= 0) {
                    do {
                        wv = status[idx];
                    } while ((wv >> 62) == 0ull);
                }
                const unsigned has_prefix = __ballot_sync(0xffffffffu, (wv >> 62) == 2ull);
                const int first = has_prefix ? __ffs(has_prefix) - 1 : 31;
                unsigned long long val = lane <= first ? (wv & LB_VALUE_MASK) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
                excl += val;
                if (has_prefix) break;
                t -= 32;
            }
        }
        if (lane == 0) {
            status[tile] = LB_FLAG_PREFIX | (excl + total);
            s_excl = excl;
            s_total = total;
        }
    }
    __syncthreads();
    unsigned long long run = s_excl + s_warp[warp] + (incl - sum);
#pragma unroll
    for (int i = 0; i < LB_ITEMS; ++i) {
        if (base + i < n) op.store(base + i, run, v[i]);
        run += v[i];
    }
    if (tile == num_tiles - 1 && threadIdx.x == 0) op.finish(s_excl + s_total);
}

}  // namespace slic
