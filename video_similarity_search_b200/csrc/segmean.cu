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

__global__ void chunk_count_kernel(const int* __restrict__ offsets, int num_clust, int* __restrict__ nchunks) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < num_clust) {
        int cnt = offsets[c + 1] - offsets[c];
        nchunks[c] = (cnt + ROWS_PER_CHUNK - 1) / ROWS_PER_CHUNK;
    }
    if (c == num_clust) nchunks[c] = 0;
}

// chunk_base: exclusive scan of nchunks over clusters, length num_clust + 1 (last = total chunks).
template <bool VEC4>
__global__ void __launch_bounds__(SM_THREADS) segmean_chunk_kernel(const float* __restrict__ data,
                                                                   const int* __restrict__ order,
                                                                   const int* __restrict__ offsets,
                                                                   const int* __restrict__ chunk_base, int num_clust,
                                                                   int d, double* __restrict__ out,
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
    const int end_all = offsets[c + 1];
    const int end = min(begin + ROWS_PER_CHUNK, end_all);
    const double cnt = (double)(end_all - offsets[c]);
    const bool direct = (n_chunks == 1);
    double* dst = direct ? out + (int64_t)c * d : partial + (int64_t)chunk * d;

    if (VEC4) {
        const int d4 = d >> 2;
        for (int k = threadIdx.x; k < d4; k += SM_THREADS) {
            double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
            int r = begin;
            // two rows in flight per iteration to keep more loads outstanding
            for (; r + 1 < end; r += 2) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(data + (int64_t)order[r] * d) + k);
                const float4 w = __ldg(reinterpret_cast<const float4*>(data + (int64_t)order[r + 1] * d) + k);
                a0 += (double)v.x; a1 += (double)v.y; a2 += (double)v.z; a3 += (double)v.w;
                a0 += (double)w.x; a1 += (double)w.y; a2 += (double)w.z; a3 += (double)w.w;
            }
            if (r < end) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(data + (int64_t)order[r] * d) + k);
                a0 += (double)v.x; a1 += (double)v.y; a2 += (double)v.z; a3 += (double)v.w;
            }
            double2 lo2, hi2;
            if (direct) {
                lo2 = make_double2(a0 / cnt, a1 / cnt);
                hi2 = make_double2(a2 / cnt, a3 / cnt);
            } else {
                lo2 = make_double2(a0, a1);
                hi2 = make_double2(a2, a3);
            }
            reinterpret_cast<double2*>(dst)[2 * k] = lo2;
            reinterpret_cast<double2*>(dst)[2 * k + 1] = hi2;
        }
    } else {
        for (int k = threadIdx.x; k < d; k += SM_THREADS) {
            double a = 0;
            for (int r = begin; r < end; ++r) a += (double)__ldg(data + (int64_t)order[r] * d + k);
            dst[k] = direct ? a / cnt : a;
        }
    }
}

// clusters that span several chunks: add the partial sums, divide by the count.  One CTA owns FIN_COLS
// columns of one cluster; its FIN_LANES thread rows each add every FIN_LANES-th chunk (fixed order), then the
// lanes are combined by a fixed-order shared-memory tree - deterministic and parallel over (cluster, column
// tile, lane) instead of one serial chain per cluster.
constexpr int FIN_COLS = 8, FIN_LANES = 32;

__global__ void __launch_bounds__(FIN_COLS * FIN_LANES) segmean_finalize_kernel(const int* __restrict__ offsets,
                                                                                const int* __restrict__ chunk_base,
                                                                                int num_clust, int d,
                                                                                const double* __restrict__ partial,
                                                                                double* __restrict__ out) {
    __shared__ double red[FIN_LANES][FIN_COLS];
    const int c = blockIdx.x;
    const int first = chunk_base[c], n_chunks = chunk_base[c + 1] - first;
    if (n_chunks <= 1) return;
    const int cx = threadIdx.x % FIN_COLS, ly = threadIdx.x / FIN_COLS;
    const double cnt = (double)(offsets[c + 1] - offsets[c]);
    for (int k0 = blockIdx.y * FIN_COLS; k0 < d; k0 += gridDim.y * FIN_COLS) {
        const int k = k0 + cx;
        double a = 0;
        if (k < d)
            for (int j = ly; j < n_chunks; j += FIN_LANES) a += partial[(int64_t)(first + j) * d + k];
        red[ly][cx] = a;
        __syncthreads();
        for (int s = FIN_LANES / 2; s > 0; s >>= 1) {
            if (ly < s) red[ly][cx] += red[ly + s][cx];
            __syncthreads();
        }
        if (ly == 0 && k < d) out[(int64_t)c * d + k] = red[0][cx] / cnt;
        __syncthreads();
    }
}

}  // namespace slic

extern "C" int slic_segmented_mean(const float* data_dev, const int32_t* labels_dev, int64_t n, int32_t d,
                                   int32_t num_clust, double* out_dev, slic_stream_t stream) {
    using namespace slic;
    SLIC_REQUIRE(n > 0 && n < ((int64_t)1 << 31) && d > 0 && num_clust > 0, "segmented_mean: bad shape");
    SLIC_REQUIRE(data_dev && labels_dev && out_dev, "segmented_mean: null pointer");
    cudaStream_t st = as_stream(stream);
    Scratch order, offsets, nchunks, chunk_base, partial;
    SLIC_CUDA_OK(order.alloc(n * sizeof(int), st));
    SLIC_CUDA_OK(offsets.alloc(((int64_t)num_clust + 1) * sizeof(int), st));
    SLIC_PROPAGATE(slic_group_by_label(labels_dev, n, num_clust, order.as<int>(), offsets.as<int>(), stream));
    SLIC_CUDA_OK(nchunks.alloc(((int64_t)num_clust + 1) * sizeof(int), st));
    SLIC_CUDA_OK(chunk_base.alloc(((int64_t)num_clust + 1) * sizeof(int), st));
    chunk_count_kernel<<<(unsigned)ceil_div((int64_t)num_clust + 1, 256), 256, 0, st>>>(offsets.as<int>(), num_clust,
                                                                                        nchunks.as<int>());
    SLIC_LAUNCH_OK();
    SLIC_PROPAGATE(exclusive_scan_i32(nchunks.as<int>(), chunk_base.as<int>(), (int64_t)num_clust + 1, nullptr, st));
    // every cluster is non-empty (labels are dense), so chunks <= n / ROWS + num_clust
    const int64_t max_chunks = n / ROWS_PER_CHUNK + num_clust;
    SLIC_CUDA_OK(partial.alloc(max_chunks * (int64_t)d * sizeof(double), st));
    const bool vec4 = (d % 4 == 0) && ((reinterpret_cast<uintptr_t>(data_dev) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(out_dev) & 15) == 0);
    if (vec4)
        segmean_chunk_kernel<true><<<(unsigned)max_chunks, SM_THREADS, 0, st>>>(
            data_dev, order.as<int>(), offsets.as<int>(), chunk_base.as<int>(), num_clust, d, out_dev,
            partial.as<double>());
    else
        segmean_chunk_kernel<false><<<(unsigned)max_chunks, SM_THREADS, 0, st>>>(
            data_dev, order.as<int>(), offsets.as<int>(), chunk_base.as<int>(), num_clust, d, out_dev,
            partial.as<double>());
    SLIC_LAUNCH_OK();
    {
        int col_tiles = (d + FIN_COLS - 1) / FIN_COLS;
        if (col_tiles > 64) col_tiles = 64;
        dim3 grid((unsigned)num_clust, (unsigned)col_tiles);
        segmean_finalize_kernel<<<grid, FIN_COLS * FIN_LANES, 0, st>>>(offsets.as<int>(), chunk_base.as<int>(), num_clust, d,
                                                                       partial.as<double>(), out_dev);
    }
    SLIC_LAUNCH_OK();
    return SLIC_OK;
}
