// Device-wide building blocks used by K2/K3/K4: int32 exclusive scan, histogram, and a stable
// LSD radix sort of (key, value) pairs.  All HBM-bound / launch-bound integer work: coalesced
// 128-bit accesses, shared-memory staging, no tensor cores.
#include "common.cuh"
#include "lookback.cuh"
#include "primitives.cuh"

namespace slic {

// ------------------------------------------------------------------------------------------
// exclusive scan: ONE launch (decoupled look-back, lookback.cuh) + one memset of its state
// ------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;

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

struct PlainScanOp {
    const int* in;
    int* out;
    int64_t n;
    int* total;
    __device__ int64_t size() const { return n; }
    __device__ unsigned long long load(int64_t i) const { return lb_pair(0, in[i]); }
    __device__ void store(int64_t i, unsigned long long excl, unsigned long long) const { out[i] = lb_b(excl); }
    __device__ void finish(unsigned long long t) const {
        if (total) *total = lb_b(t);
    }
};

int exclusive_scan_i32(const int* in, int* out, int64_t n, int* total_out, cudaStream_t st) {
    if (n <= 0) {
        if (total_out) SLIC_CUDA_OK(cudaMemsetAsync(total_out, 0, sizeof(int), st));
        return SLIC_OK;
    }
    Scratch state;
    SLIC_CUDA_OK(state.alloc(lookback_state_bytes(n), st));
    SLIC_CUDA_OK(cudaMemsetAsync(state.ptr, 0, lookback_state_bytes(n), st));
    PlainScanOp op = {in, out, n, total_out};
    lookback_scan_kernel<PlainScanOp><<<lookback_grid(n), LB_THREADS, 0, st>>>(op, state.as<unsigned long long>());
    SLIC_LAUNCH_OK();
    return SLIC_OK;
}

// ------------------------------------------------------------------------------------------
// histogram of int32 keys in [0, bins)
// ------------------------------------------------------------------------------------------
// Warp-aggregated: lanes holding the same key elect one leader that adds the group's size, so a label array
// with few distinct values (the coarse FINCH levels) does not serialise on a handful of addresses.
// Keys outside [0, bins) are not counted (the asynchronous FINCH driver enqueues this on neighbour arrays whose
// completeness it only learns later: a row the search left unsettled carries the index 0x7fffffff).
__global__ void histogram_kernel(const int* __restrict__ keys, int64_t n, int* __restrict__ counts, int64_t bins) {
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_round = (n + 31) / 32 * 32;  // keep whole warps in the loop for the match
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        const bool valid = i < n;
        const int key = valid ? keys[i] : -1 - lane;
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        if (valid && key >= 0 && key < bins && (__ffs(peers) - 1) == lane) atomicAdd(&counts[key], __popc(peers));
    }
}

int histogram_i32(const int* keys, int64_t n, int* counts, int64_t bins, cudaStream_t st) {
    SLIC_CUDA_OK(cudaMemsetAsync(counts, 0, bins * sizeof(int), st));
    if (n <= 0) return SLIC_OK;
    int blocks = (int)((n + 1023) / 1024 < (int64_t)num_sms() * 8 ? (n + 1023) / 1024 : (int64_t)num_sms() * 8);
    histogram_kernel<<<blocks, 256, 0, st>>>(keys, n, counts, bins);
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

// fallback for n >= 2^30 (the one-sweep status words hold 30-bit counts): count + table scan + scatter per pass
static int stable_sort_pairs_multi(const int* keys_in, const int* vals_in, int64_t n, int key_bits, int* keys_out,
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

// ------------------------------------------------------------------------------------------
// one-sweep variant: 1 histogram launch for all passes + 1 launch per pass
// ------------------------------------------------------------------------------------------
// Pass kernel: a CTA takes a tile ticket (4096 keys, held in registers), builds the tile's digit histogram, publishes
// it per digit (status word = flag << 30 | count), looks back over the earlier tiles' words until it meets an
// inclusive prefix (thread t follows digit t, all digits in parallel), and scatters its keys - stable, with the same
// per-round match-any ranking as rs_scatter_kernel.  The global digit bases come from the up-front histogram.
constexpr unsigned OS_AGG = 1u << 30, OS_PREFIX = 2u << 30, OS_MASK = (1u << 30) - 1u;
constexpr int OS_MAX_PASSES = 4;

template <int RB>
__global__ void __launch_bounds__(RS_THREADS) os_hist_kernel(const int* __restrict__ keys, int64_t n, int passes,
                                                             unsigned* __restrict__ ghist) {
    constexpr int R = 1 << RB;
    __shared__ unsigned sh[OS_MAX_PASSES * R];
    for (int i = threadIdx.x; i < passes * R; i += RS_THREADS) sh[i] = 0u;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * RS_THREADS;
    for (int64_t i = (int64_t)blockIdx.x * RS_THREADS + threadIdx.x; i < n; i += stride) {
        const int key = keys[i];
        for (int p = 0; p < passes; ++p) atomicAdd(&sh[p * R + ((key >> (RB * p)) & (R - 1))], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * R; i += RS_THREADS)
        if (sh[i]) atomicAdd(&ghist[i], sh[i]);
}

template <int RB>
__global__ void __launch_bounds__(RS_THREADS) os_pass_kernel(const int* __restrict__ keys_in,
                                                             const int* __restrict__ vals_in, int64_t n, int shift,
                                                             const unsigned* __restrict__ ghist_p,
                                                             unsigned* __restrict__ status_raw,
                                                             unsigned* __restrict__ ticket, int* __restrict__ keys_out,
                                                             int* __restrict__ vals_out, int rounds) {
    constexpr int R = 1 << RB;
    constexpr int PER = R / RS_THREADS;   // digits followed per thread (1 or 2)
    static_assert(R % RS_THREADS == 0, "radix must be a multiple of the block size");
    __shared__ int running[R];
    __shared__ int warp_hist[RS_WARPS][R];
    __shared__ int s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < R; i += RS_THREADS) running[i] = 0;
    if (threadIdx.x == 0) s_tile = (int)atomicAdd(ticket, 1u);
    __syncthreads();
    const int tile = s_tile;
    const int64_t base = (int64_t)tile * rounds * RS_THREADS;
    if (base >= n) return;
    volatile unsigned* status = status_raw;

#pragma unroll 4
    for (int r = 0; r < rounds; ++r) {
        const int64_t i = base + (int64_t)r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&running[(keys_in[i] >> shift) & (R - 1)], 1);
    }
    __syncthreads();
    // global base of every digit: exclusive scan of the up-front histogram (digit d = k * 256 + thread)
    int gbase[PER], tile_cnt[PER];
    int carry = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int d = k * RS_THREADS + threadIdx.x;
        int total;
        gbase[k] = carry + block_exclusive_scan((int)ghist_p[d], &total);
        carry += total;
        tile_cnt[k] = running[d];
    }
    // publish + look back, per digit
    int start[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int d = k * RS_THREADS + threadIdx.x;
        unsigned excl = 0u;
        if (tile > 0) {
            status[(int64_t)tile * R + d] = OS_AGG | (unsigned)tile_cnt[k];
            for (int t = tile - 1; t >= 0; --t) {
                unsigned w;
                do {
                    w = status[(int64_t)t * R + d];
                } while ((w >> 30) == 0u);
                excl += w & OS_MASK;
                if ((w >> 30) == 2u) break;
            }
        }
        status[(int64_t)tile * R + d] = OS_PREFIX | (excl + (unsigned)tile_cnt[k]);
        start[k] = gbase[k] + (int)excl;
    }
    __syncthreads();   // every thread has read its running[] count
#pragma unroll
    for (int k = 0; k < PER; ++k) running[k * RS_THREADS + threadIdx.x] = start[k];
    // stable scatter, 256 keys per round in index order
#pragma unroll 1
    for (int r = 0; r < rounds; ++r) {
        if (base + (int64_t)r * RS_THREADS >= n) break;
        for (int i = threadIdx.x; i < RS_WARPS * R; i += RS_THREADS) (&warp_hist[0][0])[i] = 0;
        __syncthreads();
        const int64_t i = base + (int64_t)r * RS_THREADS + threadIdx.x;
        const bool valid = i < n;
        const int kk = valid ? keys_in[i] : 0;   // second read of the tile: L1 / L2 hit
        const int digit = valid ? ((kk >> shift) & (R - 1)) : R + lane;   // invalid lanes never match anyone
        const unsigned peers = __match_any_sync(0xffffffffu, digit);
        const int rank_in_warp = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank_in_warp == 0) warp_hist[warp][digit] = __popc(peers);
        __syncthreads();
        if (valid) {
            int before = 0;
            for (int w = 0; w < warp; ++w) before += warp_hist[w][digit];
            const int pos = running[digit] + before + rank_in_warp;
            if (keys_out) keys_out[pos] = kk;
            vals_out[pos] = vals_in ? vals_in[i] : (int)i;
        }
        __syncthreads();
        for (int d = threadIdx.x; d < R; d += RS_THREADS) {
            int tot = 0;
#pragma unroll
            for (int w = 0; w < RS_WARPS; ++w) tot += warp_hist[w][d];
            running[d] += tot;
        }
        __syncthreads();   // the next round clears warp_hist: keep that after this sum
    }
}

template <int RB>
static int one_sweep_sort(const int* keys_in, const int* vals_in, int64_t n, int passes, int* keys_out, int* vals_out,
                          cudaStream_t st) {
    constexpr int R = 1 << RB;
    // keys per tile: enough tiles to put two CTAs on every SM when n allows it (the look-back chain is short either way)
    int rounds = (int)(n / ((int64_t)2 * num_sms() * RS_THREADS));
    rounds = rounds < 2 ? 2 : (rounds > RS_ROUNDS ? RS_ROUNDS : rounds);
    const int64_t tiles = ceil_div(n, (int64_t)rounds * RS_THREADS);
    // state: [passes][R] histogram, [passes] tickets (padded to 4), [passes][tiles][R] status words - one memset
    const size_t hist_words = (size_t)passes * R, ticket_words = OS_MAX_PASSES;
    const size_t status_words = (size_t)passes * tiles * R;
    Scratch state, k_tmp, v_tmp, k_tmp2;
    SLIC_CUDA_OK(state.alloc((hist_words + ticket_words + status_words) * sizeof(unsigned), st));
    SLIC_CUDA_OK(cudaMemsetAsync(state.ptr, 0, (hist_words + ticket_words + status_words) * sizeof(unsigned), st));
    unsigned* ghist = state.as<unsigned>();
    unsigned* tickets = ghist + hist_words;
    unsigned* status = tickets + ticket_words;
    if (passes > 1) {
        SLIC_CUDA_OK(k_tmp.alloc(n * sizeof(int), st));
        SLIC_CUDA_OK(v_tmp.alloc(n * sizeof(int), st));
    }
    // the caller may not want the sorted keys: intermediate passes still need them
    int* k_final = keys_out;
    if (!k_final && passes >= 3) SLIC_CUDA_OK(k_tmp2.alloc(n * sizeof(int), st));   // "out"-side passes before the last
    int hist_blocks = (int)(ceil_div(n, 4 * RS_THREADS) < (int64_t)num_sms() * 4 ? ceil_div(n, 4 * RS_THREADS) : (int64_t)num_sms() * 4);
    os_hist_kernel<RB><<<hist_blocks, RS_THREADS, 0, st>>>(keys_in, n, passes, ghist);
    SLIC_LAUNCH_OK();
    const int* src_k = keys_in;
    const int* src_v = vals_in;
    for (int p = 0; p < passes; ++p) {
        const bool last = p == passes - 1;
        const bool to_out = ((passes - 1 - p) % 2) == 0;
        int* dst_k = to_out ? (k_final ? k_final : (last ? nullptr : k_tmp2.as<int>())) : k_tmp.as<int>();
        int* dst_v = to_out ? vals_out : v_tmp.as<int>();
        os_pass_kernel<RB><<<(unsigned)tiles, RS_THREADS, 0, st>>>(src_k, src_v, n, RB * p, ghist + (size_t)p * R,
                                                                   status + (size_t)p * tiles * R, tickets + p, dst_k, dst_v,
                                                                   rounds);
        SLIC_LAUNCH_OK();
        src_k = dst_k;
        src_v = dst_v;
    }
    return SLIC_OK;
}

int stable_sort_pairs_i32(const int* keys_in, const int* vals_in, int64_t n, int key_bits, int* keys_out,
                          int* vals_out, cudaStream_t st) {
    if (n <= 0) return SLIC_OK;
    if (key_bits < 1) key_bits = 1;
    if (n >= ((int64_t)1 << 30)) return stable_sort_pairs_multi(keys_in, vals_in, n, key_bits, keys_out, vals_out, st);
    const int p8 = (key_bits + 7) / 8, p9 = (key_bits + 8) / 9;
    if (p9 < p8) return one_sweep_sort<9>(keys_in, vals_in, n, p9, keys_out, vals_out, st);   // e.g. 17-18 bits: 2 passes
    return one_sweep_sort<8>(keys_in, vals_in, n, p8, keys_out, vals_out, st);
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
