// K3: per-cluster float64 means of float32 rows (clustering/finch.py:58-71, cool_mean).
//
// The reference sorts the rows by label and differences a float64 running sum at the cluster
// boundaries; that is a float64 segmented mean (equal to <= 1e-14).  Here the rows
// are grouped by label (stable, so every cluster is summed in ascending row order - deterministic),
// each cluster is cut into chunks of ROWS_PER_CHUNK rows, and one CTA reduces one chunk with
// 128-bit coalesced row reads.  Single-chunk clusters (the common case at level 0) write their sum AND
// mean directly; multi-chunk clusters go through per-chunk partial sums that are added in chunk order.
//
// Launches per call: grouping (histogram, scan, one-sweep sort: 1 + passes) unless the caller already holds the CSR
// row pointers (the FINCH driver gets them from the components pass), then chunk scan, chunk sums, finalize.
// Every kernel takes the cluster count from DEVICE memory when asked to, so the driver can enqueue the means before
// the host has read the count back.
//
// Bound: HBM.  Algorithmic bytes per call: N*D*4 (rows) + N*4 (labels) + C*D*8 (means).
#include "common.cuh"
#include "lookback.cuh"
#include "primitives.cuh"

namespace slic {

constexpr int ROWS_PER_CHUNK = 32;
constexpr int SM_THREADS = 128;

// One scan over the clusters: chunk_base[c] = chunks of the clusters before c, pbase[c] = partial-sum rows of the
// multi-chunk clusters before c; multi-chunk clusters are appended to `multi` (any order: each is finalised alone).
// totals: [0] chunks, [1] multi-chunk clusters (atomic counter, zeroed by the caller), [2] clusters.
struct ChunkScanOp {
    const int* offsets;
    const int* num_clust_dev;   // nullptr: num_clust_host
    int num_clust_host;
    int* chunk_base;            // [C + 1]
    int* chunk_cluster;         // [chunks]: the cluster of every chunk (no search in the chunk kernel)
    int max_chunks;
    int* pbase;                 // [C]
    int* multi;                 // [<= n / 33]
    int* totals;
    // outputs, touched here only for EMPTY clusters (no chunk will ever write their rows): sum 0, count 0, mean NaN (0 / 0,
    // what the reference's division gives) - FINCH labels are dense so this is for direct callers of the C ABI
    double* sums;
    double* means;
    int* counts;
    int d;
    // (a device count above the host bound means the caller's bound was wrong: clamp - nothing is written out of
    // bounds - and the caller, who reads the count back, repeats the call with a larger bound)
    __device__ int64_t size() const {
        if (!num_clust_dev) return num_clust_host;
        const int c = *num_clust_dev;
        return c < num_clust_host ? c : num_clust_host;
    }
    __device__ unsigned long long load(int64_t c) const {
        const int cnt = offsets[c + 1] - offsets[c];
        const int nc = (cnt + ROWS_PER_CHUNK - 1) / ROWS_PER_CHUNK;
        return lb_pair(nc, nc > 1 ? nc : 0);
    }
    __device__ void store(int64_t c, unsigned long long excl, unsigned long long item) const {
        const int first = lb_a(excl), nc = lb_a(item);
        chunk_base[c] = first;
        pbase[c] = lb_b(excl);
        for (int j = 0; j < nc && first + j < max_chunks; ++j) chunk_cluster[first + j] = (int)c;
        if (lb_b(item)) multi[atomicAdd(totals + 1, 1)] = (int)c;
        if (nc == 0) {
            for (int k = 0; k < d; ++k) {
                if (sums) sums[c * d + k] = 0.0;
                if (means) means[c * d + k] = nan("");
            }
            if (counts) counts[c] = 0;
        }
    }
    __device__ void finish(unsigned long long t) const {
        const int64_t c = size();
        chunk_base[c] = lb_a(t);
        totals[0] = lb_a(t);
        totals[2] = (int)c;
    }
};

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

// Persistent over the chunks (their number lives on the device).  A single-chunk cluster gets its sum, row count and
// mean here; a chunk of a multi-chunk cluster writes a partial sum (and partial count) for the finalize kernel.
template <typename T, bool VEC4>
__global__ void __launch_bounds__(SM_THREADS) segsum_chunk_kernel(const T* __restrict__ data, const int* __restrict__ weights,
                                                                  const int* __restrict__ order,
                                                                  const int* __restrict__ offsets,
                                                                  const int* __restrict__ chunk_base,
                                                                  const int* __restrict__ chunk_cluster,
                                                                  const int* __restrict__ pbase,
                                                                  const int* __restrict__ totals, int d,
                                                                  double* __restrict__ sums, int* __restrict__ counts,
                                                                  double* __restrict__ means, double* __restrict__ partial,
                                                                  int* __restrict__ partial_cnt) {
    const int total_chunks = totals[0];
    for (int chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
        const int c = __ldg(chunk_cluster + chunk);
        const int first_chunk = chunk_base[c];
        const int n_chunks = chunk_base[c + 1] - first_chunk;
        const int begin = offsets[c] + (chunk - first_chunk) * ROWS_PER_CHUNK;
        const int end = min(begin + ROWS_PER_CHUNK, offsets[c + 1]);
        const bool single = n_chunks == 1;
        const int64_t prow = single ? 0 : (int64_t)pbase[c] + (chunk - first_chunk);
        double* dst = single ? sums + (int64_t)c * d : partial + prow * d;
        int cnt = end - begin;
        if (weights) {   // rows are clusters of the finer partition: their row counts add up (uniform across the CTA)
            cnt = 0;
            for (int r = begin; r < end; ++r) cnt += __ldg(weights + order[r]);
        }
        const double rows = (double)cnt;
        if (VEC4) {
            const int d4 = d >> 2;
            for (int k = threadIdx.x; k < d4; k += SM_THREADS) {
                Acc4 a = {0, 0, 0, 0};
                int r = begin;
                for (; r + 1 < end; r += 2) {   // two rows in flight per iteration
                    const T* r0 = data + (int64_t)order[r] * d;
                    const T* r1 = data + (int64_t)order[r + 1] * d;
                    add_row(a, r0, k);
                    add_row(a, r1, k);
                }
                if (r < end) add_row(a, data + (int64_t)order[r] * d, k);
                reinterpret_cast<double2*>(dst)[2 * k] = make_double2(a.a0, a.a1);
                reinterpret_cast<double2*>(dst)[2 * k + 1] = make_double2(a.a2, a.a3);
                if (single && means) {
                    double2* m = reinterpret_cast<double2*>(means + (int64_t)c * d);
                    m[2 * k] = make_double2(a.a0 / rows, a.a1 / rows);
                    m[2 * k + 1] = make_double2(a.a2 / rows, a.a3 / rows);
                }
            }
        } else {
            for (int k = threadIdx.x; k < d; k += SM_THREADS) {
                double a = 0;
                for (int r = begin; r < end; ++r) a += (double)__ldg(data + (int64_t)order[r] * d + k);
                dst[k] = a;
                if (single && means) means[(int64_t)c * d + k] = a / rows;
            }
        }
        if (threadIdx.x == 0) {
            if (single) {
                if (counts) counts[c] = cnt;
            } else {
                partial_cnt[prow] = cnt;
            }
        }
    }
}

// clusters that span several chunks: add the partial sums.  One CTA per such cluster (grid-stride); for every tile of
// FIN_COLS columns its FIN_LANES thread rows each add every FIN_LANES-th chunk (fixed order), then the lanes are
// combined by a fixed-order shared-memory tree - deterministic whatever the grid.
constexpr int FIN_COLS = 32, FIN_LANES = 8, FIN_DIRECT_MAX = 32;

__global__ void __launch_bounds__(FIN_COLS * FIN_LANES) segsum_finalize_kernel(const int* __restrict__ multi,
                                                                               const int* __restrict__ totals,
                                                                               const int* __restrict__ chunk_base,
                                                                               const int* __restrict__ pbase, int d,
                                                                               const double* __restrict__ partial,
                                                                               const int* __restrict__ partial_cnt,
                                                                               double* __restrict__ sums,
                                                                               int* __restrict__ counts,
                                                                               double* __restrict__ means) {
    __shared__ double red[FIN_LANES][FIN_COLS];
    __shared__ int s_cnt;
    const int n_multi = totals[1];
    const int cx = threadIdx.x % FIN_COLS, ly = threadIdx.x / FIN_COLS;
    for (int m = blockIdx.x; m < n_multi; m += gridDim.x) {
        const int c = multi[m];
        const int n_chunks = chunk_base[c + 1] - chunk_base[c];
        const int64_t first = pbase[c];
        if (threadIdx.x < 32) {   // integer sum: order-free
            int cnt = 0;
            for (int j = threadIdx.x; j < n_chunks; j += 32) cnt += partial_cnt[first + j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            if (threadIdx.x == 0) s_cnt = cnt;
        }
        __syncthreads();
        const double rows = (double)s_cnt;
        if (n_chunks <= FIN_DIRECT_MAX) {
            // few chunks (the usual case: a cluster of a few dozen rows): every thread owns columns and adds the chunks
            // in order - coalesced, no barriers.  (Which path runs depends on n_chunks only: deterministic.)
            for (int k = threadIdx.x; k < d; k += FIN_COLS * FIN_LANES) {
                double a = 0;
                for (int j = 0; j < n_chunks; ++j) a += partial[(first + j) * d + k];
                sums[(int64_t)c * d + k] = a;
                if (means) means[(int64_t)c * d + k] = a / rows;
            }
        } else
        for (int k0 = 0; k0 < d; k0 += FIN_COLS) {
            const int k = k0 + cx;
            double a = 0;
            if (k < d)
                for (int j = ly; j < n_chunks; j += FIN_LANES) a += partial[(first + j) * d + k];
            red[ly][cx] = a;
            __syncthreads();
            for (int s = FIN_LANES / 2; s > 0; s >>= 1) {
                if (ly < s) red[ly][cx] += red[ly + s][cx];
                __syncthreads();
            }
            if (ly == 0 && k < d) {
                sums[(int64_t)c * d + k] = red[0][cx];
                if (means) means[(int64_t)c * d + k] = red[0][cx] / rows;
            }
            __syncthreads();   // red is rewritten by the next tile
        }
        if (threadIdx.x == 0 && counts) counts[c] = s_cnt;
        __syncthreads();   // s_cnt is rewritten by the next trip
    }
}

// Per-cluster float64 sums of the rows of `data` (float32 originals, or float64 sums of a finer partition) from the
// CSR grouping (order = rows sorted by cluster, stable; offsets = row pointers), rows of a cluster added in ascending
// row order; optional counts (weighted by `weights` when given) and means.  num_clust_dev (optional): the cluster
// count on the device - num_clust is then only an upper bound used to size temporaries and grids.
template <typename T>
int cluster_sums_csr(const T* data, const int* weights, const int* order, const int* offsets, int64_t n, int d,
                     int num_clust, const int* num_clust_dev, double* sums_out, int* counts_out, double* means_out,
                     cudaStream_t st) {
    Scratch chunk_base, chunk_cluster, pbase, multi, totals, partial, partial_cnt, state;
    const int64_t max_chunks = n / ROWS_PER_CHUNK + num_clust;
    SLIC_CUDA_OK(chunk_base.alloc(((int64_t)num_clust + 1) * sizeof(int), st));
    SLIC_CUDA_OK(chunk_cluster.alloc((max_chunks + 1) * sizeof(int), st));
    SLIC_CUDA_OK(pbase.alloc(((int64_t)num_clust + 1) * sizeof(int), st));
    // a multi-chunk cluster has > ROWS_PER_CHUNK rows: at most n / 33 of them, holding at most n / 32 + n / 33 chunks
    const int64_t max_multi = n / (ROWS_PER_CHUNK + 1) < num_clust ? n / (ROWS_PER_CHUNK + 1) : num_clust;
    const int64_t max_partial = n / ROWS_PER_CHUNK + max_multi;
    SLIC_CUDA_OK(multi.alloc((max_multi + 1) * sizeof(int), st));
    SLIC_CUDA_OK(partial.alloc((max_partial + 1) * (int64_t)d * sizeof(double), st));
    SLIC_CUDA_OK(partial_cnt.alloc((max_partial + 1) * sizeof(int), st));
    // totals[4] followed by the scan state: one memset
    const size_t state_bytes = 4 * sizeof(int) * 2 + lookback_state_bytes(num_clust);
    SLIC_CUDA_OK(state.alloc(state_bytes, st));
    SLIC_CUDA_OK(cudaMemsetAsync(state.ptr, 0, state_bytes, st));
    int* tot = state.as<int>();
    unsigned long long* scan_state = reinterpret_cast<unsigned long long*>(state.as<int>() + 8);
    ChunkScanOp op = {offsets, num_clust_dev, num_clust, chunk_base.as<int>(), chunk_cluster.as<int>(), (int)max_chunks,
                      pbase.as<int>(), multi.as<int>(), tot, sums_out, means_out, counts_out, d};
    lookback_scan_kernel<ChunkScanOp><<<lookback_grid(num_clust), LB_THREADS, 0, st>>>(op, scan_state);
    SLIC_LAUNCH_OK();
    const bool vec4 = (d % 4 == 0) && ((reinterpret_cast<uintptr_t>(data) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(sums_out) & 15) == 0) &&
                      (!means_out || (reinterpret_cast<uintptr_t>(means_out) & 15) == 0);
    const int64_t want = (int64_t)num_sms() * 16;
    const unsigned grid = (unsigned)(max_chunks < want ? (max_chunks > 0 ? max_chunks : 1) : want);
    if (vec4)
        segsum_chunk_kernel<T, true><<<grid, SM_THREADS, 0, st>>>(data, weights, order, offsets, chunk_base.as<int>(),
                                                                  chunk_cluster.as<int>(), pbase.as<int>(), tot, d, sums_out,
                                                                  counts_out, means_out,
                                                                  partial.as<double>(), partial_cnt.as<int>());
    else
        segsum_chunk_kernel<T, false><<<grid, SM_THREADS, 0, st>>>(data, weights, order, offsets, chunk_base.as<int>(),
                                                                   chunk_cluster.as<int>(), pbase.as<int>(), tot, d, sums_out,
                                                                   counts_out, means_out,
                                                                   partial.as<double>(), partial_cnt.as<int>());
    SLIC_LAUNCH_OK();
    if (max_multi > 0) {
        const int64_t fin_want = (int64_t)num_sms() * 8;
        const unsigned fgrid = (unsigned)(max_multi < fin_want ? max_multi : fin_want);
        segsum_finalize_kernel<<<fgrid, FIN_COLS * FIN_LANES, 0, st>>>(multi.as<int>(), tot, chunk_base.as<int>(),
                                                                      pbase.as<int>(), d, partial.as<double>(),
                                                                      partial_cnt.as<int>(), sums_out, counts_out, means_out);
        SLIC_LAUNCH_OK();
    }
    return SLIC_OK;
}
template int cluster_sums_csr<float>(const float*, const int*, const int*, const int*, int64_t, int, int, const int*,
                                     double*, int*, double*, cudaStream_t);
template int cluster_sums_csr<double>(const double*, const int*, const int*, const int*, int64_t, int, int, const int*,
                                      double*, int*, double*, cudaStream_t);

static int bits_for_labels(int64_t num_labels) {
    int b = 1;
    while (((int64_t)1 << b) < num_labels && b < 31) ++b;
    return b;
}

// rows sorted by label (stable) when the CSR row pointers are already known (the components pass emits them)
int order_rows_by_label(const int* labels, int64_t n, int64_t num_labels_bound, int* order, cudaStream_t st) {
    return stable_sort_pairs_i32(labels, nullptr, n, bits_for_labels(num_labels_bound), nullptr, order, st);
}

// labels -> grouping -> sums: the public entry points (labels dense in [0, num_clust), count known on the host)
template <typename T>
static int cluster_sums_impl(const T* data, const int* weights, const int* labels, int64_t n, int d, int num_clust,
                             double* sums_out, int* counts_out, double* means_out, cudaStream_t st) {
    Scratch order, offsets, sums_tmp;
    SLIC_CUDA_OK(order.alloc(n * sizeof(int), st));
    SLIC_CUDA_OK(offsets.alloc(((int64_t)num_clust + 1) * sizeof(int), st));
    SLIC_PROPAGATE(slic_group_by_label(labels, n, num_clust, order.as<int>(), offsets.as<int>(), st));
    if (!sums_out) {
        SLIC_CUDA_OK(sums_tmp.alloc((int64_t)num_clust * d * sizeof(double), st));
        sums_out = sums_tmp.as<double>();
    }
    return cluster_sums_csr<T>(data, weights, order.as<int>(), offsets.as<int>(), n, d, num_clust, nullptr, sums_out,
                               counts_out, means_out, st);
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
