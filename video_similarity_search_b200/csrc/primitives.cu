// Device-wide building blocks used by K2/K3/K4: int32 exclusive scan, histogram, and a stable
// LSD radix sort of (key, value) pairs.  All HBM-bound / launch-bound integer work: coalesced
// 128-bit accesses, shared-memory staging, no tensor cores.
#include "common.cuh"
#include "primitives.cuh"

namespace slic {

// ------------------------------------------------------------------------------------------
// exclusive scan
// ------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
    // 256 threads; returns exclusive prefix of v over the block, *total = block sum.
    __shared__ int warp_sums[SCAN_THREADS / 32];
    __shared__ int block_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
        int wi = w;
#pragma unroll
        for (int o = 1; o < SCAN_THREADS / 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = wi - w;
        if (lane == SCAN_THREADS / 32 - 1) block_total = wi;
    }
    __syncthreads();
    int res = warp_sums[warp] + incl - v;
    *total = block_total;
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tiles_kernel(const int* __restrict__ in, int* __restrict__ out,
                                                                  int* __restrict__ tile_sums, int64_t n) {
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int sum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        sum += v[i];
    }
    int total;
    int pre = block_exclusive_scan(sum, &total);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = pre;
        pre += v[i];
    }
    if (threadIdx.x == 0 && tile_sums) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_add_kernel(int* __restrict__ out, const int* __restrict__ tile_pre,
                                                                int64_t n) {
    const int add = tile_pre[blockIdx.x];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
        if (base + i < n) out[base + i] += add;
}

__global__ void write_total_kernel(const int* __restrict__ in, const int* __restrict__ out, int64_t n, int* total) {
    *total = n > 0 ? out[n - 1] + in[n - 1] : 0;
}

int exclusive_scan_i32(const int* in, int* out, int64_t n, int* total_out, cudaStream_t st) {
    if (n <= 0) {
        if (total_out) SLIC_CUDA_OK(cudaMemsetAsync(total_out, 0, sizeof(int), st));
        return SLIC_OK;
    }
    const int64_t tiles = ceil_div(n, SCAN_TILE);
    if (tiles == 1) {
        scan_tiles_kernel<<<1, SCAN_THREADS, 0, st>>>(in, out, nullptr, n);
        SLIC_LAUNCH_OK();
    } else {
        Scratch sums, sums_scanned;
        SLIC_CUDA_OK(sums.alloc(tiles * sizeof(int), st));
        SLIC_CUDA_OK(sums_scanned.alloc(tiles * sizeof(int), st));
        scan_tiles_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, out, sums.as<int>(), n);
        SLIC_LAUNCH_OK();
        SLIC_PROPAGATE(exclusive_scan_i32(sums.as<int>(), sums_scanned.as<int>(), tiles, nullptr, st));
        scan_add_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(out, sums_scanned.as<int>(), n);
        SLIC_LAUNCH_OK();
    }
    if (total_out) {
        write_total_kernel<<<1, 1, 0, st>>>(in, out, n, total_out);
        SLIC_LAUNCH_OK();
    }
    return SLIC_OK;
}

// ------------------------------------------------------------------------------------------
// histogram of int32 keys in [0, bins)
// ------------------------------------------------------------------------------------------
// Warp-aggregated: lanes holding the same key elect one leader that adds the group's size, so a label array
// with few distinct values (the coarse FINCH levels) does not serialise on a handful of addresses.
__global__ void histogram_kernel(const int* __restrict__ keys, int64_t n, int* __restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_round = (n + 31) / 32 * 32;  // keep whole warps in the loop for the match
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        const bool valid = i < n;
        const int key = valid ? keys[i] : -1 - lane;
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        if (valid && (__ffs(peers) - 1) == lane) atomicAdd(&counts[key], __popc(peers));
    }
}

int histogram_i32(const int* keys, int64_t n, int* counts, int64_t bins, cudaStream_t st) {
    SLIC_CUDA_OK(cudaMemsetAsync(counts, 0, bins * sizeof(int), st));
    if (n <= 0) return SLIC_OK;
    int blocks = (int)((n + 1023) / 1024 < (int64_t)num_sms() * 8 ? (n + 1023) / 1024 : (int64_t)num_sms() * 8);
    histogram_kernel<<<blocks, 256, 0, st>>>(keys, n, counts);
    SLIC_LAUNCH_OK();
    return SLIC_OK;
}

// ------------------------------------------------------------------------------------------
// stable LSD radix sort of (int32 key >= 0, int32 value), 8 bits per pass
// ------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ROUNDS = 16;                         // rounds of 256 keys per block
constexpr int RS_TILE = RS_THREADS * RS_ROUNDS;      // 4096 keys per block

// pass 1: per-block digit counts, table laid out [digit][block] so that one exclusive scan of the
// table yields the global base of every (digit, block) bucket.
__global__ void __launch_bounds__(RS_THREADS) rs_count_kernel(const int* __restrict__ keys, int64_t n, int shift,
                                                              int* __restrict__ table, int num_blocks) {
    __shared__ int hist[256];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int r = 0; r < RS_ROUNDS; ++r) {
        int64_t i = base + r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&hist[(keys[i] >> shift) & 255], 1);
    }
    __syncthreads();
    table[(int64_t)threadIdx.x * num_blocks + blockIdx.x] = hist[threadIdx.x];
}

// pass 2: stable scatter.  Inside a block keys are taken in index order, 256 at a time; inside a
// round the rank of a key among equal digits is (#equal in earlier warps) + (#equal in lower lanes).
__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const int* __restrict__ keys_in,
                                                                const int* __restrict__ vals_in, int64_t n, int shift,
                                                                const int* __restrict__ table_scanned, int num_blocks,
                                                                int* __restrict__ keys_out, int* __restrict__ vals_out) {
    __shared__ int running[256];              // next free global slot per digit for this block
    __shared__ int warp_hist[RS_WARPS][256];  // per-round per-warp digit counts
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    running[threadIdx.x] = table_scanned[(int64_t)threadIdx.x * num_blocks + blockIdx.x];
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
    for (int r = 0; r < RS_ROUNDS; ++r) {
        if (base + (int64_t)r * RS_THREADS >= n) break;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) warp_hist[w][threadIdx.x] = 0;
        __syncthreads();
        const int64_t i = base + (int64_t)r * RS_THREADS + threadIdx.x;
        const bool valid = i < n;
        int key = 0, val = 0, digit = 256 + lane;  // invalid lanes never match anyone
        if (valid) {
            key = keys_in[i];
            val = vals_in ? vals_in[i] : (int)i;
            digit = (key >> shift) & 255;
        }
        const unsigned peers = __match_any_sync(0xffffffffu, digit);
        const int rank_in_warp = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank_in_warp == 0) warp_hist[warp][digit] = __popc(peers);
        __syncthreads();
        if (valid) {
            int before = 0;
            for (int w = 0; w < warp; ++w) before += warp_hist[w][digit];
            const int pos = running[digit] + before + rank_in_warp;
            keys_out[pos] = key;
            vals_out[pos] = val;
        }
        __syncthreads();
        int tot = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) tot += warp_hist[w][threadIdx.x];
        running[threadIdx.x] += tot;
        __syncthreads();  // the next round clears warp_hist: keep that after this sum
    }
}

int stable_sort_pairs_i32(const int* keys_in, const int* vals_in, int64_t n, int key_bits, int* keys_out,
                          int* vals_out, cudaStream_t st) {
    if (n <= 0) return SLIC_OK;
    int passes = (key_bits + 7) / 8;
    if (passes < 1) passes = 1;
    const int num_blocks = (int)ceil_div(n, RS_TILE);
    const int64_t tab = (int64_t)256 * num_blocks;
    Scratch table, table_scanned, k_tmp, v_tmp;
    SLIC_CUDA_OK(table.alloc(tab * sizeof(int), st));
    SLIC_CUDA_OK(table_scanned.alloc(tab * sizeof(int), st));
    SLIC_CUDA_OK(k_tmp.alloc(n * sizeof(int), st));
    SLIC_CUDA_OK(v_tmp.alloc(n * sizeof(int), st));
    // ping-pong so that the last pass lands in the caller's buffers
    const int* src_k = keys_in;
    const int* src_v = vals_in;
    for (int p = 0; p < passes; ++p) {
        const bool to_out = ((passes - 1 - p) % 2) == 0;
        int* dst_k = to_out ? keys_out : k_tmp.as<int>();
        int* dst_v = to_out ? vals_out : v_tmp.as<int>();
        rs_count_kernel<<<num_blocks, RS_THREADS, 0, st>>>(src_k, n, 8 * p, table.as<int>(), num_blocks);
        SLIC_LAUNCH_OK();
        SLIC_PROPAGATE(exclusive_scan_i32(table.as<int>(), table_scanned.as<int>(), tab, nullptr, st));
        rs_scatter_kernel<<<num_blocks, RS_THREADS, 0, st>>>(src_k, src_v, n, 8 * p, table_scanned.as<int>(),
                                                            num_blocks, dst_k, dst_v);
        SLIC_LAUNCH_OK();
        src_k = dst_k;
        src_v = dst_v;
    }
    return SLIC_OK;
}

__global__ void iota_kernel(int* out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (int)i;
}

int iota_i32(int* out, int64_t n, cudaStream_t st) {
    if (n <= 0) return SLIC_OK;
    iota_kernel<<<(unsigned)(ceil_div(n, 256) < 4096 ? ceil_div(n, 256) : 4096), 256, 0, st>>>(out, n);
    SLIC_LAUNCH_OK();
    return SLIC_OK;
}

}  // namespace slic
