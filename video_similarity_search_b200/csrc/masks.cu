// K4a/K4b: label-equality masks over FINCH labels, and the hit@k reduction of the retrieval scripts.
//
// Reference call sites: models/infoNCE.py:281-283 (UberNCE  B x (1+K) positive mask),
// loss/triplet_loss.py:136-142, 254-261, 291-297 (per-label positive / negative index masks),
// evaluate.py:287-307 and iic_retrieve_clips.py:298-306 (hit@k).
//
// Bound: HBM writes.  Byte mask: 8 (na + nb) + na (nb + prepend) bytes; bit mask: na * ceil(nb/32) * 4.
// Each thread produces 16 output bytes (one 128-bit store, coalesced across the warp); the column
// labels of the 16 columns come through the read-only path and are reused across the rows a CTA walks.
#include "common.cuh"
#include "primitives.cuh"

namespace slic {

constexpr int MASK_THREADS = 256;
constexpr int MASK_ROWS_PER_CTA = 16;

// out row stride = nb + prepend.  A CTA covers MASK_THREADS*16 output columns x MASK_ROWS_PER_CTA rows.
__global__ void __launch_bounds__(MASK_THREADS) label_mask_u8_kernel(const int64_t* __restrict__ a, int64_t na,
                                                                     const int64_t* __restrict__ b, int64_t nb,
                                                                     int prepend, int negate,
                                                                     uint8_t* __restrict__ out) {
    const int64_t stride = nb + prepend;
    const int64_t col0 = ((int64_t)blockIdx.x * MASK_THREADS + threadIdx.x) * 16;  // output column
    if (col0 >= stride) return;
    // labels of my 16 output columns (output column c maps to b[c - prepend]; c < prepend is the ones column)
    int64_t lab[16];
    unsigned valid = 0, ones = 0;
#pragma unroll
    for (int t = 0; t < 16; ++t) {
        const int64_t c = col0 + t;
        lab[t] = 0;
        if (c < stride) {
            valid |= 1u << t;
            if (c < prepend) ones |= 1u << t; else lab[t] = __ldg(b + (c - prepend));
        }
    }
    const uint8_t neg = negate ? 1 : 0;
    const int64_t row0 = (int64_t)blockIdx.y * MASK_ROWS_PER_CTA;
    for (int r = 0; r < MASK_ROWS_PER_CTA; ++r) {
        const int64_t row = row0 + r;
        if (row >= na) break;
        const int64_t la = __ldg(a + row);
        uint8_t bytes[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            uint8_t eq = (lab[t] == la) ? 1 : 0;
            bytes[t] = ((ones >> t) & 1u) ? (uint8_t)1 : (uint8_t)(eq ^ neg);
        }
        uint8_t* dst = out + row * stride + col0;
        if (valid == 0xffffu && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
            uint4 v;
            v.x = bytes[0] | (bytes[1] << 8) | (bytes[2] << 16) | ((unsigned)bytes[3] << 24);
            v.y = bytes[4] | (bytes[5] << 8) | (bytes[6] << 16) | ((unsigned)bytes[7] << 24);
            v.z = bytes[8] | (bytes[9] << 8) | (bytes[10] << 16) | ((unsigned)bytes[11] << 24);
            v.w = bytes[12] | (bytes[13] << 8) | (bytes[14] << 16) | ((unsigned)bytes[15] << 24);
            *reinterpret_cast<uint4*>(dst) = v;
        } else {
#pragma unroll
            for (int t = 0; t < 16; ++t)
                if ((valid >> t) & 1u) dst[t] = bytes[t];
        }
    }
}

// bit-packed: one warp ballot = 32 columns of one row.  A warp owns 32 consecutive columns and walks rows.
__global__ void __launch_bounds__(MASK_THREADS) label_mask_bits_kernel(const int64_t* __restrict__ a, int64_t na,
                                                                       const int64_t* __restrict__ b, int64_t nb,
                                                                       int negate, uint32_t* __restrict__ out,
                                                                       int64_t words_per_row) {
    const int lane = threadIdx.x & 31;
    const int64_t word = (int64_t)blockIdx.x * (MASK_THREADS / 32) + (threadIdx.x >> 5);
    if (word >= words_per_row) return;
    const int64_t col = word * 32 + lane;
    const bool in_range = col < nb;
    const int64_t lb = in_range ? __ldg(b + col) : 0;
    const int64_t row0 = (int64_t)blockIdx.y * 32;
    uint32_t mine = 0;  // lane r keeps the word of row row0 + r: b is read once per 32 rows, one store per lane
    for (int r = 0; r < 32; ++r) {
        const int64_t row = row0 + r;
        if (row >= na) break;
        const bool eq = in_range && ((lb == __ldg(a + row)) != (negate != 0));
        const uint32_t bits = __ballot_sync(0xffffffffu, eq);
        if (lane == r) mine = bits;
    }
    const int64_t row = row0 + lane;
    if (row < na) out[row * words_per_row + word] = mine;
}

// hits[m] += 1 for every query whose label appears among the labels of its first ks[m] neighbours
__global__ void __launch_bounds__(256) hit_at_k_kernel(const int* __restrict__ topk, int64_t nq, int k_stride,
                                                       const int64_t* __restrict__ q_labels,
                                                       const int64_t* __restrict__ x_labels,
                                                       const int* __restrict__ ks, int num_ks, int* hits) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= nq) return;
    const int64_t lq = q_labels[q];
    const int kmax = ks[num_ks - 1];
    int first = 0x7fffffff;  // rank of the first neighbour carrying the query's label
    for (int base = 0; base < kmax; base += 32) {
        const int r = base + lane;
        const bool hit = r < kmax && x_labels[topk[q * k_stride + r]] == lq;
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (m) {
            first = base + __ffs(m) - 1;
            break;
        }
    }
    if (lane < num_ks && first < ks[lane]) atomicAdd(&hits[lane], 1);
}

// online_train.py:648-652 - `out[idxs[i]] = labels[i]` for i = 0..n-1 in order: where the distributed sampler padded
// the epoch with repeated dataset indices the LAST occurrence wins.  Pass 1 elects, per output slot, the largest source
// position (integer atomic max: exact, order-free); pass 2 copies that source's value.
__global__ void scatter_elect_kernel(const int64_t* __restrict__ pos, int64_t n, int64_t n_out, int* __restrict__ winner,
                                     int* __restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t p = pos[i];
    if (p < 0 || p >= n_out) {
        atomicAdd(bad, 1);
        return;
    }
    atomicMax(winner + p, (int)i);
}
__global__ void scatter_copy_kernel(const int* __restrict__ values, const int* __restrict__ winner, int64_t n_out, int fill,
                                    int* __restrict__ out) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_out) return;
    const int w = winner[p];
    out[p] = w >= 0 ? values[w] : fill;
}

// ---- dense relabelling: np.unique(labels, return_inverse=True) for int32 labels ------------------------------
__global__ void bias_keys_kernel(const int* __restrict__ labels, int64_t n, int* __restrict__ keys) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = labels[i] ^ (int)0x80000000;   // signed order -> order of the bit pattern (what the radix passes see)
}
__global__ void flag_boundaries_kernel(const int* __restrict__ sorted_keys, int64_t n, int* __restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = (i == 0 || sorted_keys[i] != sorted_keys[i - 1]) ? 1 : 0;
}
__global__ void assign_dense_ids_kernel(const int* __restrict__ sorted_keys, const int* __restrict__ order,
                                        const int* __restrict__ flags, const int* __restrict__ before, int64_t n,
                                        int* __restrict__ dense, int* __restrict__ uniq) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int id = before[i] + flags[i] - 1;   // boundaries at or before position i, minus one
    dense[order[i]] = id;
    if (uniq && flags[i]) uniq[id] = sorted_keys[i] ^ (int)0x80000000;
}

}  // namespace slic

extern "C" {

int slic_dense_labels(const int32_t* labels_dev, int64_t n, int32_t* dense_out_dev, int32_t* uniq_out_dev,
                      int32_t* num_out_dev, slic_stream_t stream) {
    using namespace slic;
    SLIC_REQUIRE(n >= 1 && n < ((int64_t)1 << 31), "dense_labels: bad shape");
    SLIC_REQUIRE(labels_dev && dense_out_dev && num_out_dev, "dense_labels: null pointer");
    cudaStream_t st = as_stream(stream);
    Scratch keys, sorted, order, flags, before;
    SLIC_CUDA_OK(keys.alloc((size_t)n * sizeof(int), st));
    SLIC_CUDA_OK(sorted.alloc((size_t)n * sizeof(int), st));
    SLIC_CUDA_OK(order.alloc((size_t)n * sizeof(int), st));
    SLIC_CUDA_OK(flags.alloc((size_t)n * sizeof(int), st));
    SLIC_CUDA_OK(before.alloc((size_t)n * sizeof(int), st));
    const unsigned blocks = (unsigned)ceil_div(n, 256);
    bias_keys_kernel<<<blocks, 256, 0, st>>>(labels_dev, n, keys.as<int>());
    SLIC_LAUNCH_OK();
    SLIC_PROPAGATE(stable_sort_pairs_i32(keys.as<int>(), nullptr, n, 32, sorted.as<int>(), order.as<int>(), st));
    flag_boundaries_kernel<<<blocks, 256, 0, st>>>(sorted.as<int>(), n, flags.as<int>());
    SLIC_LAUNCH_OK();
    SLIC_PROPAGATE(exclusive_scan_i32(flags.as<int>(), before.as<int>(), n, num_out_dev, st));
    assign_dense_ids_kernel<<<blocks, 256, 0, st>>>(sorted.as<int>(), order.as<int>(), flags.as<int>(), before.as<int>(), n,
                                                    dense_out_dev, uniq_out_dev);
    SLIC_LAUNCH_OK();
    return SLIC_OK;
}

int slic_scatter_last_wins(const int32_t* values_dev, const int64_t* positions_dev, int64_t n, int64_t n_out,
                           int32_t fill, int32_t* out_dev, int32_t* out_of_range_dev, slic_stream_t stream) {
    using namespace slic;
    SLIC_REQUIRE(n >= 0 && n < ((int64_t)1 << 31) && n_out >= 0, "scatter_last_wins: bad shape");
    SLIC_REQUIRE((n == 0 || (values_dev && positions_dev)) && (n_out == 0 || out_dev) && out_of_range_dev,
                 "scatter_last_wins: null pointer");
    cudaStream_t st = as_stream(stream);
    SLIC_CUDA_OK(cudaMemsetAsync(out_of_range_dev, 0, sizeof(int), st));
    if (n_out == 0) return SLIC_OK;
    Scratch winner;
    SLIC_CUDA_OK(winner.alloc((size_t)n_out * sizeof(int), st));
    SLIC_CUDA_OK(cudaMemsetAsync(winner.ptr, 0xff, (size_t)n_out * sizeof(int), st));   // -1: nobody wrote this slot
    if (n > 0) {
        scatter_elect_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(positions_dev, n, n_out, winner.as<int>(),
                                                                         out_of_range_dev);
        SLIC_LAUNCH_OK();
    }
    scatter_copy_kernel<<<(unsigned)ceil_div(n_out, 256), 256, 0, st>>>(values_dev, winner.as<int>(), n_out, fill, out_dev);
    SLIC_LAUNCH_OK();
    return SLIC_OK;
}

int slic_label_mask_u8(const int64_t* a_dev, int64_t na, const int64_t* b_dev, int64_t nb, int32_t prepend_ones,
                       int32_t negate, uint8_t* out_dev, slic_stream_t stream) {
    SLIC_REQUIRE(na >= 0 && nb >= 0 && a_dev && b_dev && out_dev, "label_mask_u8: bad arguments");
    const int prepend = prepend_ones ? 1 : 0;
    if (na == 0 || nb + prepend == 0) return SLIC_OK;
    dim3 grid((unsigned)slic::ceil_div(nb + prepend, slic::MASK_THREADS * 16),
              (unsigned)slic::ceil_div(na, slic::MASK_ROWS_PER_CTA));
    SLIC_REQUIRE(grid.y <= 65535, "label_mask_u8: too many rows for one launch");
    slic::label_mask_u8_kernel<<<grid, slic::MASK_THREADS, 0, slic::as_stream(stream)>>>(a_dev, na, b_dev, nb, prepend,
                                                                                        negate, out_dev);
    SLIC_LAUNCH_OK();
    return SLIC_OK;
}

int slic_label_mask_bits(const int64_t* a_dev, int64_t na, const int64_t* b_dev, int64_t nb, int32_t negate,
                         uint32_t* out_dev, slic_stream_t stream) {
    SLIC_REQUIRE(na >= 0 && nb >= 0 && a_dev && b_dev && out_dev, "label_mask_bits: bad arguments");
    if (na == 0 || nb == 0) return SLIC_OK;
    const int64_t words = slic::ceil_div(nb, 32);
    dim3 grid((unsigned)slic::ceil_div(words, slic::MASK_THREADS / 32), (unsigned)slic::ceil_div(na, 32));
    SLIC_REQUIRE(grid.y <= 65535, "label_mask_bits: too many rows for one launch");
    slic::label_mask_bits_kernel<<<grid, slic::MASK_THREADS, 0, slic::as_stream(stream)>>>(a_dev, na, b_dev, nb, negate,
                                                                                          out_dev, words);
    SLIC_LAUNCH_OK();
    return SLIC_OK;
}

int slic_hit_at_k(const int32_t* topk_idx_dev, int64_t nq, int32_t k_stride, const int64_t* q_labels_dev,
                  const int64_t* x_labels_dev, const int32_t* ks_dev, int32_t num_ks, int32_t* hits_out_dev,
                  slic_stream_t stream) {
    SLIC_REQUIRE(nq >= 0 && k_stride > 0 && num_ks > 0 && num_ks <= 32, "hit_at_k: bad shape (at most 32 ks)");
    SLIC_REQUIRE(topk_idx_dev && q_labels_dev && x_labels_dev && ks_dev && hits_out_dev, "hit_at_k: null pointer");
    cudaStream_t st = slic::as_stream(stream);
    SLIC_CUDA_OK(cudaMemsetAsync(hits_out_dev, 0, num_ks * sizeof(int), st));
    if (nq == 0) return SLIC_OK;
    slic::hit_at_k_kernel<<<(unsigned)slic::ceil_div(nq, 8), 256, 0, st>>>(topk_idx_dev, nq, k_stride, q_labels_dev,
                                                                           x_labels_dev, ks_dev, num_ks, hits_out_dev);
    SLIC_LAUNCH_OK();
    return SLIC_OK;
}

}  // extern "C"
