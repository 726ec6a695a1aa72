// K3: per-cluster float64 means of float32 rows (clustering/finch.py:58-71, cool_mean).
//
// The reference sorts the rows by label and differences a float64 running sum at the cluster
// boundaries; that is a float64 segmented mean (equal to <= 1e-14).  Here the rows
// are grouped by label (stable, so every cluster is summed in ascending row order - deterministic),
// each cluster is cut into chunks of ROWS_PER_CHUNK rows, and one CTA reduces one chunk with
// 128-bit coalesced row reads.  Single-chunk clusters (the common case at level 0) write their mean
// directly; multi-chunk clusters go through per-chunk partial sums that are added in chunk order.
//
// Bound: HBM.  Algorithmic bytes per call: N*D*4 (rows) + N*4 (labels) + C*D*8 (means).
#include "common.cuh"
#include "primitives.cuh"

namespace slic {

constexpr int ROWS_PER_CHUNK = 32;
constexpr int SM_THREADS = 128;

// multi[0 .. *n_multi) lists the clusters that span more than one chunk (any order)
__global__ void chunk_count_kernel(const int* __restrict__ offsets, int num_clust, int* __restrict__ nchunks,
                                   int* __restrict__ multi, int* __restrict__ n_multi) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < num_clust) {
        int cnt = offsets[c + 1] - offsets[c];
        const int nc = (cnt + ROWS_PER_CHUNK - 1) / ROWS_PER_CHUNK;
        nchunks[c] = nc;
        if (nc > 1) multi[atomicAdd(n_multi, 1)] = c;
    }
    if (c == num_clust) nchunks[c] = 0;
}

// 128-bit row loads of either input type, widened to float64
struct Acc4 { double a0, a1, a2, a3; };
__device__ __forceinline__ void add_row(Acc4& a, const float* __restrict__ row, int k) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(row) + k);
    a.a0 += (double)v.x; a.a1 += (double)v.y; a.a2 += (double)v.z; a.a3 += (double)v.w;
}
__device__ __forceinline__ void add_row(Acc4& a, const double* __restrict__ row, int k) {
    const double2 v = __ldg(reinterpret_cast<const double2*>(row) + 2 * k);
    const double2 w = __ldg(reinterpret_cast<const double2*>(row) + 2 * k + 1);
    a.a0 += v.x; a.a1 += v.y; a.a2 += w.x; a.a3 += w.y;
}

// chunk_base: exclusive scan of nchunks over clusters, length num_clust + 1 (last = total chunks).
// Writes SUMS: a single-chunk cluster straight into `sums`, a multi-chunk cluster into `partial`.
template <typename T, bool VEC4>
__global__ void __launch_bounds__(SM_THREADS) segsum_chunk_kernel(const T* __restrict__ data,
                                                                  const int* __restrict__ order,
                                                                  const int* __restrict__ offsets,
                                                                  const int* __restrict__ chunk_base, int num_clust,
                                                                  int d, double* __restrict__ sums,
                                                                  double* __restrict__ partial) {
    __shared__ int s_cluster;
    const int chunk = blockIdx.x;
    if (chunk >= chunk_base[num_clust]) return;
    if (threadIdx.x == 0) {
        // largest c with chunk_base[c] <= chunk
        int lo = 0, hi = num_clust;  // invariant: chunk_base[lo] <= chunk < chunk_base[hi]
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (chunk_base[mid] <= chunk) lo = mid; else hi = mid;
        }
        s_cluster = lo;
    }
    __syncthreads();
    const int c = s_cluster;
    const int first_chunk = chunk_base[c];
    const int n_chunks = chunk_base[c + 1] - first_chunk;
    const int begin = offsets[c] + (chunk - first_chunk) * ROWS_PER_CHUNK;
    const int end = min(begin + ROWS_PER_CHUNK, offsets[c + 1]);
    double* dst = (n_chunks == 1) ? sums + (int64_t)c * d : partial + (int64_t)chunk * d;

    if (VEC4) {
        const int d4 = d >> 2;
        for (int k = threadIdx.x; k < d4; k += SM_THREADS) {
            Acc4 a = {0, 0, 0, 0};
            int r = begin;
            // two rows in flight per iteration to keep more loads outstanding
            for (; r + 1 < end; r += 2) {
                const T* r0 = data + (int64_t)order[r] * d;
                const T* r1 = data + (int64_t)order[r + 1] * d;
                add_row(a, r0, k);
                add_row(a, r1, k);
            }
            if (r < end) add_row(a, data + (int64_t)order[r] * d, k);
            reinterpret_cast<double2*>(dst)[2 * k] = make_double2(a.a0, a.a1);
            reinterpret_cast<double2*>(dst)[2 * k + 1] = make_double2(a.a2, a.a3);
        }
    } else {
        for (int k = threadIdx.x; k < d; k += SM_THREADS) {
            double a = 0;
            for (int r = begin; r < end; ++r) a += (double)__ldg(data + (int64_t)order[r] * d + k);
            dst[k] = a;
        }
    }
}

// clusters that span several chunks: add the partial sums.  One CTA owns FIN_COLS columns of one such cluster;
// its FIN_LANES thread rows each add every FIN_LANES-th chunk (fixed order), then the lanes are combined by a
// fixed-order shared-memory tree - deterministic, and parallel over (cluster, column tile, lane).
constexpr int FIN_COLS = 32, FIN_LANES = 8;

__global__ void __launch_bounds__(FIN_COLS * FIN_LANES) segsum_finalize_kernel(const int* __restrict__ multi,
                                                                               const int* __restrict__ n_multi,
                                                                               const int* __restrict__ chunk_base,
                                                                               int d, const double* __restrict__ partial,
                                                                               double* __restrict__ out) {
    __shared__ double red[FIN_LANES][FIN_COLS];
    if ((int)blockIdx.x >= *n_multi) return;
    const int c = multi[blockIdx.x];
    const int first = chunk_base[c], n_chunks = chunk_base[c + 1] - first;
    const int cx = threadIdx.x % FIN_COLS, ly = threadIdx.x / FIN_COLS;
    const int k = blockIdx.y * FIN_COLS + cx;
    double a = 0;
    if (k < d)
        for (int j = ly; j < n_chunks; j += FIN_LANES) a += partial[(int64_t)(first + j) * d + k];
    red[ly][cx] = a;
    __syncthreads();
    for (int s = FIN_LANES / 2; s > 0; s >>= 1) {
        if (ly < s) red[ly][cx] += red[ly + s][cx];
        __syncthreads();
    }
    if (ly == 0 && k < d) out[(int64_t)c * d + k] = red[0][cx];
}

// counts[c]: rows of the cluster (weights == nullptr) or the sum of its members' weights (merging sums of sums)
__global__ void cluster_counts_kernel(const int* __restrict__ order, const int* __restrict__ offsets,
                                      const int* __restrict__ weights, int num_clust, int* __restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= num_clust) return;
    const int b = offsets[c], e = offsets[c + 1];
    int cnt = e - b;
    if (weights) {
        cnt = 0;
        for (int r = b + lane; r < e; r += 32) cnt += weights[order[r]];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if (lane == 0) counts[c] = cnt;
}

__global__ void means_from_sums_kernel(const double* __restrict__ sums, const int* __restrict__ counts, int64_t total,
                                       int d, double* __restrict__ means) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) means[i] = sums[i] / (double)counts[i / d];
}

// Per-cluster float64 sums of the rows of `data` (float32 originals, or float64 sums of a finer partition),
// rows of a cluster added in ascending row order; optional counts (weighted by `weights` when given) and means.
template <typename T>
static int cluster_sums_impl(const T* data, const int* weights, const int* labels, int64_t n, int d, int num_clust,
                             double* sums_out, int* counts_out, double* means_out, cudaStream_t st) {
    Scratch order, offsets, nchunks, chunk_base, partial, counts_tmp, sums_tmp, multi;
    SLIC_CUDA_OK(order.alloc(n * sizeof(int), st));
    SLIC_CUDA_OK(offsets.alloc(((int64_t)num_clust + 1) * sizeof(int), st));
    SLIC_PROPAGATE(slic_group_by_label(labels, n, num_clust, order.as<int>(), offsets.as<int>(), st));
    SLIC_CUDA_OK(nchunks.alloc(((int64_t)num_clust + 1) * sizeof(int), st));
    SLIC_CUDA_OK(chunk_base.alloc(((int64_t)num_clust + 1) * sizeof(int), st));
    // at most n / ROWS_PER_CHUNK clusters can span several chunks; slot 0 of `multi` is the counter
    const int64_t max_multi = n / ROWS_PER_CHUNK < num_clust ? n / ROWS_PER_CHUNK : num_clust;
    SLIC_CUDA_OK(multi.alloc((max_multi + 1) * sizeof(int), st));
    SLIC_CUDA_OK(cudaMemsetAsync(multi.ptr, 0, sizeof(int), st));
    chunk_count_kernel<<<(unsigned)ceil_div((int64_t)num_clust + 1, 256), 256, 0, st>>>(
        offsets.as<int>(), num_clust, nchunks.as<int>(), multi.as<int>() + 1, multi.as<int>());
    SLIC_LAUNCH_OK();
    SLIC_PROPAGATE(exclusive_scan_i32(nchunks.as<int>(), chunk_base.as<int>(), (int64_t)num_clust + 1, nullptr, st));
    if (!sums_out) {
        SLIC_CUDA_OK(sums_tmp.alloc((int64_t)num_clust * d * sizeof(double), st));
        sums_out = sums_tmp.as<double>();
    }
    // every cluster is non-empty (labels are dense), so chunks <= n / ROWS + num_clust
    const int64_t max_chunks = n / ROWS_PER_CHUNK + num_clust;
    SLIC_CUDA_OK(partial.alloc(max_chunks * (int64_t)d * sizeof(double), st));
    const bool vec4 = (d % 4 == 0) && ((reinterpret_cast<uintptr_t>(data) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(sums_out) & 15) == 0);
    if (vec4)
        segsum_chunk_kernel<T, true><<<(unsigned)max_chunks, SM_THREADS, 0, st>>>(
            data, order.as<int>(), offsets.as<int>(), chunk_base.as<int>(), num_clust, d, sums_out, partial.as<double>());
    else
        segsum_chunk_kernel<T, false><<<(unsigned)max_chunks, SM_THREADS, 0, st>>>(
            data, order.as<int>(), offsets.as<int>(), chunk_base.as<int>(), num_clust, d, sums_out, partial.as<double>());
    SLIC_LAUNCH_OK();
    if (max_multi > 0) {
        dim3 grid((unsigned)max_multi, (unsigned)ceil_div(d, FIN_COLS));
        segsum_finalize_kernel<<<grid, FIN_COLS * FIN_LANES, 0, st>>>(multi.as<int>() + 1, multi.as<int>(),
                                                                      chunk_base.as<int>(), d, partial.as<double>(),
                                                                      sums_out);
        SLIC_LAUNCH_OK();
    }
    if (counts_out || means_out) {
        if (!counts_out) {
            SLIC_CUDA_OK(counts_tmp.alloc((int64_t)num_clust * sizeof(int), st));
            counts_out = counts_tmp.as<int>();
        }
        cluster_counts_kernel<<<(unsigned)ceil_div(num_clust, 8), 256, 0, st>>>(order.as<int>(), offsets.as<int>(), weights,
                                                                                num_clust, counts_out);
        SLIC_LAUNCH_OK();
    }
    if (means_out) {
        const int64_t total = (int64_t)num_clust * d;
        means_from_sums_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(sums_out, counts_out, total, d, means_out);
        SLIC_LAUNCH_OK();
    }
    return SLIC_OK;
}

}  // namespace slic

extern "C" int slic_segmented_mean(const float* data_dev, const int32_t* labels_dev, int64_t n, int32_t d,
                                   int32_t num_clust, double* out_dev, slic_stream_t stream) {
    using namespace slic;
    SLIC_REQUIRE(n > 0 && n < ((int64_t)1 << 31) && d > 0 && num_clust > 0, "segmented_mean: bad shape");
    SLIC_REQUIRE(data_dev && labels_dev && out_dev, "segmented_mean: null pointer");
    return cluster_sums_impl<float>(data_dev, nullptr, labels_dev, n, d, num_clust, nullptr, nullptr, out_dev,
                                    as_stream(stream));
}

extern "C" int slic_cluster_sums(const float* data_dev, const int32_t* labels_dev, int64_t n, int32_t d,
                                 int32_t num_clust, double* sums_out_dev, int32_t* counts_out_dev, double* means_out_dev,
                                 slic_stream_t stream) {
    using namespace slic;
    SLIC_REQUIRE(n > 0 && n < ((int64_t)1 << 31) && d > 0 && num_clust > 0, "cluster_sums: bad shape");
    SLIC_REQUIRE(data_dev && labels_dev && sums_out_dev && counts_out_dev, "cluster_sums: null pointer");
    return cluster_sums_impl<float>(data_dev, nullptr, labels_dev, n, d, num_clust, sums_out_dev, counts_out_dev,
                                    means_out_dev, as_stream(stream));
}

extern "C" int slic_merge_cluster_sums(const double* sums_prev_dev, const int32_t* counts_prev_dev,
                                       const int32_t* labels_dev, int64_t n_prev, int32_t d, int32_t num_clust,
                                       double* sums_out_dev, int32_t* counts_out_dev, double* means_out_dev,
                                       slic_stream_t stream) {
    using namespace slic;
    SLIC_REQUIRE(n_prev > 0 && n_prev < ((int64_t)1 << 31) && d > 0 && num_clust > 0, "merge_cluster_sums: bad shape");
    SLIC_REQUIRE(sums_prev_dev && counts_prev_dev && labels_dev && sums_out_dev && counts_out_dev,
                 "merge_cluster_sums: null pointer");
    return cluster_sums_impl<double>(sums_prev_dev, counts_prev_dev, labels_dev, n_prev, d, num_clust, sums_out_dev,
                                     counts_out_dev, means_out_dev, as_stream(stream));
}
