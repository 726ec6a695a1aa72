// K2 + K4c: connected components of the first-neighbour graph, the min_sim reduction, and the
// CSR grouping of rows by label.
//
// The reference builds the sparse link graph A = (P + I)(P + I)^T (clustering/finch.py:41-46) and
// hands it to scipy's connected_components (finch.py:54).  A has an entry for (i, nn[i]) (2 when the
// pair is mutual) and for every pair of rows sharing a first neighbour.  Here the graph is never
// built: links are enumerated straight from nn[] (direct links) and from the rows grouped by their
// first neighbour (sibling pairs), and merged with a lock-free union-find whose roots are always the
// smallest member index - so the final numbering (rank of the root among roots) is scipy's numbering.
//
// Bound: latency / HBM.  Algorithmic bytes ~ 8 n (nn in, labels out); sibling-pair distances are the
// only floating-point work and touch 2 rows per pair.
#include "common.cuh"
#include "lookback.cuh"
#include "primitives.cuh"

namespace slic {

// ---- lock-free union-find, parent[x] <= x always ------------------------------------------
__device__ __forceinline__ int uf_find(int* parent, int x) {
    volatile int* p = parent;
    int px = p[x];
    while (px != x) {
        int ppx = p[px];
        if (ppx != px) p[x] = ppx;  // path halving; only ever lowers a non-root's pointer
        x = px;
        px = ppx;
    }
    return x;
}

__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        int hi = a > b ? a : b, lo = a > b ? b : a;
        if (atomicCAS(&parent[hi], hi, lo) == hi) return;
    }
}

__global__ void cc_init_kernel(int* parent, int* size, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        parent[i] = (int)i;
        if (size) size[i] = 0;
    }
}

// direct links i - nn[i]; with the filter a link survives iff w * d <= min_sim (finch.py:51-52)
template <typename T>
__global__ void cc_union_direct_kernel(const int* __restrict__ nn, int64_t n, const T* __restrict__ dist_nn,
                                       int use_filter, double min_sim, const float* __restrict__ min_sim_dev, int* parent,
                                       int* __restrict__ bad) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int j = nn[i];
    if (j < 0 || j >= n) {   // the reference's sparse matrix constructor raises on such an index
        if (bad) atomicAdd(bad, 1);
        return;
    }
    if (j == i) return;
    if (use_filter) {
        if (min_sim_dev) min_sim = (double)*min_sim_dev;   // the float32 value of finch.py:144, still on the device
        double w = (nn[j] == (int)i) ? 2.0 : 1.0;
        if ((double)dist_nn[i] * w > min_sim) return;
    }
    uf_union(parent, (int)i, j);
}

// sibling pairs: rows are grouped by first neighbour (order / key-sorted), one warp per row i handles
// the pairs (i, j) with j later in the same group.  MODE 0: union when d <= min_sim.  MODE 1: running
// maximum of d (for min_sim itself).  MODE 2: minimum of d over all pairs (bits of the float64 distance,
// monotone for d >= 0).  MODE 3: smallest packed (i << 32 | j) among the pairs whose distance equals *aux_in.
template <typename T, int MODE>
__global__ void __launch_bounds__(256) sibling_pairs_kernel(const int* __restrict__ order,
                                                            const int* __restrict__ sorted_nn,
                                                            const int* __restrict__ offsets, int64_t n,
                                                            const T* __restrict__ unit, int d, double min_sim,
                                                            int* parent, unsigned int* max_bits,
                                                            unsigned long long* aux_out = nullptr,
                                                            const unsigned long long* aux_in = nullptr,
                                                            const float* __restrict__ min_sim_dev = nullptr) {
    const int lane = threadIdx.x & 31;
    if (MODE == 0 && min_sim_dev) min_sim = (double)*min_sim_dev;
    const int64_t p = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= n) return;
    const int hub = sorted_nn[p];
    if (hub < 0 || hub >= n) return;
    const int end = offsets[hub + 1];
    if (p + 1 >= end) return;
    const int i = order[p];
    const T* ri = unit + (int64_t)i * d;
    float local_max = 0.f;
    for (int q = (int)p + 1; q < end; ++q) {
        const int j = order[q];
        const double s = warp_dot<T>(ri, unit + (int64_t)j * d, d, lane);
        const T dist = cosine_distance_from_sim<T>(s);
        if (MODE == 0) {
            if (lane == 0 && (double)dist <= min_sim) uf_union(parent, i, j);
        } else if (MODE == 1) {
            local_max = fmaxf(local_max, (float)dist);
        } else if (MODE == 2) {
            if (lane == 0) atomicMin(aux_out, (unsigned long long)__double_as_longlong((double)dist));
        } else {
            if (lane == 0 && (unsigned long long)__double_as_longlong((double)dist) == *aux_in) {
                const unsigned long long lo = i < j ? i : j, hi = i < j ? j : i;
                atomicMin(aux_out, (lo << 32) | hi);
            }
        }
    }
    if (MODE == 1 && lane == 0) atomicMax(max_bits, __float_as_uint(local_max));
}

// min_sim contribution of the direct links: max_i w_i * d(i, nn[i]) in float32 (finch.py:144)
template <typename T>
__global__ void direct_max_kernel(const int* __restrict__ nn, int64_t n, const T* __restrict__ dist_nn,
                                  unsigned int* max_bits) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float v = 0.f;
    if (i < n) {
        int j = nn[i];
        if (j >= 0 && j < n && j != i) {
            float w = (nn[j] == (int)i) ? 2.f : 1.f;
            v = (float)dist_nn[i] * w;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0 && v > 0.f) atomicMax(max_bits, __float_as_uint(v));
}

// closest-link search over the direct links (finch.py:85-94): PASS 0 = minimum distance, PASS 1 = the pair
template <typename T, int PASS>
__global__ void direct_min_kernel(const int* __restrict__ nn, int64_t n, const T* __restrict__ dist_nn,
                                  unsigned long long* aux_out, const unsigned long long* aux_in) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int j = nn[i];
    if (j < 0 || j >= n || j == i) return;
    const unsigned long long bits = (unsigned long long)__double_as_longlong((double)dist_nn[i]);
    if (PASS == 0) {
        atomicMin(aux_out, bits);
    } else if (bits == *aux_in) {
        const unsigned long long lo = i < j ? i : j, hi = i < j ? j : i;
        atomicMin(aux_out, (lo << 32) | hi);
    }
}

__global__ void unpack_pair_kernel(const unsigned long long* packed, int* pair_out) {
    pair_out[0] = (int)(*packed >> 32);
    pair_out[1] = (int)(*packed & 0xffffffffull);
}

// root[i] = smallest member of i's component; size[root] (optional) = members of the component
__global__ void cc_flatten_kernel(int* parent, int64_t n, int* root, int* size) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < n;
    int r = -1 - lane;
    if (valid) {
        r = uf_find(parent, (int)i);
        root[i] = r;
    }
    if (size) {   // warp-aggregated: neighbouring rows often share a root
        const unsigned peers = __match_any_sync(0xffffffffu, r);
        if (valid && (__ffs(peers) - 1) == lane) atomicAdd(&size[r], __popc(peers));
    }
}

// ONE scan over the rows yields, at every root, its rank among the roots (= scipy's component number: components are
// numbered by smallest member) and the exclusive sum of the earlier roots' sizes (= the CSR offset of its cluster)
struct RootScanOp {
    const int* root;
    const int* csize;   // nullptr: ranks only
    int64_t n;
    int* rank;
    int* off;
    int* num_clust;
    __device__ int64_t size() const { return n; }
    __device__ unsigned long long load(int64_t i) const {
        const bool is_root = root[i] == (int)i;
        return lb_pair(is_root ? 1 : 0, (is_root && csize) ? csize[i] : 0);
    }
    __device__ void store(int64_t i, unsigned long long excl, unsigned long long) const {
        rank[i] = lb_a(excl);
        if (off) off[i] = lb_b(excl);
    }
    __device__ void finish(unsigned long long t) const { *num_clust = lb_a(t); }
};

// labels[i] = rank of i's root; with csr outputs every root also writes its cluster's CSR offset and row count
__global__ void cc_relabel_kernel(const int* __restrict__ root, const int* __restrict__ rank, int64_t n,
                                  int* __restrict__ labels, const int* __restrict__ off, const int* __restrict__ size,
                                  const int* __restrict__ num_clust, int* __restrict__ csr_offsets,
                                  int* __restrict__ csr_counts) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int r = root[i];
    const int lab = rank[r];
    labels[i] = lab;
    if (csr_offsets) {
        if (r == (int)i) {
            csr_offsets[lab] = off[i];
            if (csr_counts) csr_counts[lab] = size[i];
        }
        if (i == 0) csr_offsets[*num_clust] = (int)n;
    }
}

__global__ void negate_count_if_bad_kernel(const int* __restrict__ bad, int* __restrict__ count) {
    if (*bad > 0) *count = -*bad;
}

__global__ void compose_kernel(const int* __restrict__ prev, const int* __restrict__ u, int64_t n,
                               int* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = prev ? u[prev[i]] : u[i];
}

static int bits_for(int64_t num_labels) {
    int b = 1;
    while (((int64_t)1 << b) < num_labels && b < 31) ++b;
    return b;
}

// order = stable argsort(labels), sorted = labels[order], offsets = CSR row pointers
static int group_rows(const int* labels, int64_t n, int64_t num_labels, int* order, int* sorted, int* offsets,
                      cudaStream_t st) {
    Scratch counts;
    SLIC_CUDA_OK(counts.alloc((num_labels + 1) * sizeof(int), st));
    SLIC_PROPAGATE(histogram_i32(labels, n, counts.as<int>(), num_labels + 1, st));
    SLIC_PROPAGATE(exclusive_scan_i32(counts.as<int>(), offsets, num_labels + 1, nullptr, st));
    SLIC_PROPAGATE(stable_sort_pairs_i32(labels, nullptr, n, bits_for(num_labels), sorted, order, st));
    return SLIC_OK;
}

template <typename T>
static int components_impl(const int* nn, int64_t n, int use_filter, double min_sim, const float* min_sim_dev,
                           const T* unit, int d, const T* dist_nn, int* labels, int* num_clust, int* csr_offsets,
                           int* csr_counts, cudaStream_t st, int* bad = nullptr) {
    const unsigned blocks = (unsigned)ceil_div(n, 256);
    const bool csr = csr_offsets != nullptr;
    Scratch parent, root, rank, size, off, state;
    SLIC_CUDA_OK(parent.alloc(n * sizeof(int), st));
    SLIC_CUDA_OK(root.alloc(n * sizeof(int), st));
    SLIC_CUDA_OK(rank.alloc(n * sizeof(int), st));
    if (csr) {
        SLIC_CUDA_OK(size.alloc(n * sizeof(int), st));
        SLIC_CUDA_OK(off.alloc(n * sizeof(int), st));
    }
    SLIC_CUDA_OK(state.alloc(lookback_state_bytes(n), st));
    SLIC_CUDA_OK(cudaMemsetAsync(state.ptr, 0, lookback_state_bytes(n), st));
    cc_init_kernel<<<blocks, 256, 0, st>>>(parent.as<int>(), csr ? size.as<int>() : nullptr, n);
    SLIC_LAUNCH_OK();
    cc_union_direct_kernel<T><<<blocks, 256, 0, st>>>(nn, n, dist_nn, use_filter, min_sim, min_sim_dev, parent.as<int>(), bad);
    SLIC_LAUNCH_OK();
    if (use_filter) {
        // without the filter every sibling pair is already joined through its hub
        Scratch order, sorted, offsets;
        SLIC_CUDA_OK(order.alloc(n * sizeof(int), st));
        SLIC_CUDA_OK(sorted.alloc(n * sizeof(int), st));
        SLIC_CUDA_OK(offsets.alloc((n + 1) * sizeof(int), st));
        SLIC_PROPAGATE(group_rows(nn, n, n, order.as<int>(), sorted.as<int>(), offsets.as<int>(), st));
        sibling_pairs_kernel<T, 0><<<(unsigned)ceil_div(n, 8), 256, 0, st>>>(
            order.as<int>(), sorted.as<int>(), offsets.as<int>(), n, unit, d, min_sim, parent.as<int>(), nullptr, nullptr,
            nullptr, min_sim_dev);
        SLIC_LAUNCH_OK();
    }
    cc_flatten_kernel<<<blocks, 256, 0, st>>>(parent.as<int>(), n, root.as<int>(), csr ? size.as<int>() : nullptr);
    SLIC_LAUNCH_OK();
    RootScanOp op = {root.as<int>(), csr ? size.as<int>() : nullptr, n, rank.as<int>(), csr ? off.as<int>() : nullptr,
                     num_clust};
    lookback_scan_kernel<RootScanOp><<<lookback_grid(n), LB_THREADS, 0, st>>>(op, state.as<unsigned long long>());
    SLIC_LAUNCH_OK();
    cc_relabel_kernel<<<blocks, 256, 0, st>>>(root.as<int>(), rank.as<int>(), n, labels, off.as<int>(), size.as<int>(),
                                              num_clust, csr_offsets, csr_counts);
    SLIC_LAUNCH_OK();
    return SLIC_OK;
}

// driver-internal form (finch_driver.cu): min_sim may still live on the device, and the CSR row pointers / row counts
// of the clusters (what the means need) come out of the same scan.  csr_offsets: room for (clusters + 1) ints.
int finch_components_csr(const int* nn, int64_t n, int use_filter, const float* min_sim_dev, const void* unit, int d,
                         int dtype, const void* dist_nn, int* labels, int* num_clust_dev, int* csr_offsets,
                         int* csr_counts, cudaStream_t st, int* bad_count_dev) {
    if (use_filter && dtype == SLIC_F64)
        return components_impl<double>(nn, n, 1, 0.0, min_sim_dev, (const double*)unit, d, (const double*)dist_nn, labels,
                                       num_clust_dev, csr_offsets, csr_counts, st, bad_count_dev);
    return components_impl<float>(nn, n, use_filter, 0.0, min_sim_dev, (const float*)unit, d, (const float*)dist_nn, labels,
                                  num_clust_dev, csr_offsets, csr_counts, st, bad_count_dev);
}

template <typename T>
static int min_sim_impl(const int* nn, int64_t n, const T* unit, int d, const T* dist_nn, float* out,
                        cudaStream_t st) {
    SLIC_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(float), st));
    direct_max_kernel<T><<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(nn, n, dist_nn, (unsigned int*)out);
    SLIC_LAUNCH_OK();
    Scratch order, sorted, offsets;
    SLIC_CUDA_OK(order.alloc(n * sizeof(int), st));
    SLIC_CUDA_OK(sorted.alloc(n * sizeof(int), st));
    SLIC_CUDA_OK(offsets.alloc((n + 1) * sizeof(int), st));
    SLIC_PROPAGATE(group_rows(nn, n, n, order.as<int>(), sorted.as<int>(), offsets.as<int>(), st));
    sibling_pairs_kernel<T, 1><<<(unsigned)ceil_div(n, 8), 256, 0, st>>>(
        order.as<int>(), sorted.as<int>(), offsets.as<int>(), n, unit, d, 0.0, nullptr, (unsigned int*)out);
    SLIC_LAUNCH_OK();
    return SLIC_OK;
}

template <typename T>
static int closest_link_impl(const int* nn, int64_t n, const T* unit, int d, const T* dist_nn, int* pair_out,
                             cudaStream_t st) {
    Scratch aux, order, sorted, offsets;
    SLIC_CUDA_OK(aux.alloc(2 * sizeof(unsigned long long), st));
    SLIC_CUDA_OK(cudaMemsetAsync(aux.ptr, 0xff, 2 * sizeof(unsigned long long), st));
    unsigned long long* a = aux.as<unsigned long long>();
    SLIC_CUDA_OK(order.alloc(n * sizeof(int), st));
    SLIC_CUDA_OK(sorted.alloc(n * sizeof(int), st));
    SLIC_CUDA_OK(offsets.alloc((n + 1) * sizeof(int), st));
    SLIC_PROPAGATE(group_rows(nn, n, n, order.as<int>(), sorted.as<int>(), offsets.as<int>(), st));
    const unsigned b1 = (unsigned)ceil_div(n, 256), b8 = (unsigned)ceil_div(n, 8);
    direct_min_kernel<T, 0><<<b1, 256, 0, st>>>(nn, n, dist_nn, a, nullptr);
    SLIC_LAUNCH_OK();
    sibling_pairs_kernel<T, 2><<<b8, 256, 0, st>>>(order.as<int>(), sorted.as<int>(), offsets.as<int>(), n, unit, d, 0.0,
                                                  nullptr, nullptr, a, nullptr);
    SLIC_LAUNCH_OK();
    direct_min_kernel<T, 1><<<b1, 256, 0, st>>>(nn, n, dist_nn, a + 1, a);
    SLIC_LAUNCH_OK();
    sibling_pairs_kernel<T, 3><<<b8, 256, 0, st>>>(order.as<int>(), sorted.as<int>(), offsets.as<int>(), n, unit, d, 0.0,
                                                  nullptr, nullptr, a + 1, a);
    SLIC_LAUNCH_OK();
    unpack_pair_kernel<<<1, 1, 0, st>>>(a + 1, pair_out);
    SLIC_LAUNCH_OK();
    return SLIC_OK;
}

}  // namespace slic

extern "C" {

int slic_finch_closest_link(const int32_t* nn_dev, int64_t n, const void* unit_dev, int32_t d, int32_t dtype,
                            const void* dist_nn_dev, int32_t* pair_out_dev, slic_stream_t stream) {
    SLIC_REQUIRE(n > 1 && n < ((int64_t)1 << 31), "finch_closest_link: n out of range");
    SLIC_REQUIRE(nn_dev && unit_dev && dist_nn_dev && pair_out_dev && d > 0, "finch_closest_link: null pointer");
    SLIC_REQUIRE(dtype == SLIC_F32 || dtype == SLIC_F64, "finch_closest_link: bad dtype");
    cudaStream_t st = slic::as_stream(stream);
    if (dtype == SLIC_F64)
        return slic::closest_link_impl<double>(nn_dev, n, (const double*)unit_dev, d, (const double*)dist_nn_dev,
                                               pair_out_dev, st);
    return slic::closest_link_impl<float>(nn_dev, n, (const float*)unit_dev, d, (const float*)dist_nn_dev, pair_out_dev,
                                          st);
}

int slic_finch_components(const int32_t* nn_dev, int64_t n, int32_t use_filter, double min_sim, const void* unit_dev,
                          int32_t d, int32_t dtype, const void* dist_nn_dev, int32_t* labels_out_dev,
                          int32_t* num_clust_out_dev, slic_stream_t stream) {
    SLIC_REQUIRE(n >= 0 && n < ((int64_t)1 << 31), "finch_components: n out of range");
    SLIC_REQUIRE(nn_dev && labels_out_dev && num_clust_out_dev, "finch_components: null pointer");
    if (use_filter) {
        SLIC_REQUIRE(unit_dev && dist_nn_dev && d > 0, "finch_components: the min_sim filter needs unit rows and distances");
        SLIC_REQUIRE(dtype == SLIC_F32 || dtype == SLIC_F64, "finch_components: bad dtype");
    }
    cudaStream_t st = slic::as_stream(stream);
    if (n == 0) {
        SLIC_CUDA_OK(cudaMemsetAsync(num_clust_out_dev, 0, sizeof(int), st));
        return SLIC_OK;
    }
    // neighbour indices outside [0, n) are never followed (no out-of-bounds access); the reference's sparse-matrix
    // constructor raises on them (finch.py:41-43) - here the count comes back NEGATED (-number of such rows)
    slic::Scratch bad;
    SLIC_CUDA_OK(bad.alloc(sizeof(int), st));
    SLIC_CUDA_OK(cudaMemsetAsync(bad.ptr, 0, sizeof(int), st));
    int status;
    if (use_filter && dtype == SLIC_F64)
        status = slic::components_impl<double>(nn_dev, n, 1, min_sim, nullptr, (const double*)unit_dev, d,
                                               (const double*)dist_nn_dev, labels_out_dev, num_clust_out_dev, nullptr,
                                               nullptr, st, bad.as<int>());
    else
        status = slic::components_impl<float>(nn_dev, n, use_filter, min_sim, nullptr, (const float*)unit_dev, d,
                                              (const float*)dist_nn_dev, labels_out_dev, num_clust_out_dev, nullptr,
                                              nullptr, st, bad.as<int>());
    SLIC_PROPAGATE(status);
    slic::negate_count_if_bad_kernel<<<1, 1, 0, st>>>(bad.as<int>(), num_clust_out_dev);
    SLIC_LAUNCH_OK();
    return SLIC_OK;
}

int slic_finch_min_sim(const int32_t* nn_dev, int64_t n, const void* unit_dev, int32_t d, int32_t dtype,
                       const void* dist_nn_dev, float* min_sim_out_dev, slic_stream_t stream) {
    SLIC_REQUIRE(n > 0 && n < ((int64_t)1 << 31), "finch_min_sim: n out of range");
    SLIC_REQUIRE(nn_dev && unit_dev && dist_nn_dev && min_sim_out_dev && d > 0, "finch_min_sim: null pointer");
    SLIC_REQUIRE(dtype == SLIC_F32 || dtype == SLIC_F64, "finch_min_sim: bad dtype");
    cudaStream_t st = slic::as_stream(stream);
    if (dtype == SLIC_F64)
        return slic::min_sim_impl<double>(nn_dev, n, (const double*)unit_dev, d, (const double*)dist_nn_dev,
                                          min_sim_out_dev, st);
    return slic::min_sim_impl<float>(nn_dev, n, (const float*)unit_dev, d, (const float*)dist_nn_dev,
                                     min_sim_out_dev, st);
}

int slic_compose_labels(const int32_t* prev_dev, const int32_t* u_dev, int64_t n, int32_t* out_dev,
                        slic_stream_t stream) {
    SLIC_REQUIRE(n >= 0 && u_dev && out_dev, "compose_labels: bad arguments");
    if (n == 0) return SLIC_OK;
    slic::compose_kernel<<<(unsigned)slic::ceil_div(n, 256), 256, 0, slic::as_stream(stream)>>>(prev_dev, u_dev, n,
                                                                                               out_dev);
    SLIC_LAUNCH_OK();
    return SLIC_OK;
}

int slic_group_by_label(const int32_t* labels_dev, int64_t n, int32_t num_labels, int32_t* order_out_dev,
                        int32_t* offsets_out_dev, slic_stream_t stream) {
    SLIC_REQUIRE(n >= 0 && n < ((int64_t)1 << 31) && num_labels > 0, "group_by_label: bad shape");
    SLIC_REQUIRE(labels_dev && order_out_dev && offsets_out_dev, "group_by_label: null pointer");
    cudaStream_t st = slic::as_stream(stream);
    slic::Scratch sorted;
    SLIC_CUDA_OK(sorted.alloc((n > 0 ? n : 1) * sizeof(int), st));
    return slic::group_rows(labels_dev, n, num_labels, order_out_dev, sorted.as<int>(), offsets_out_dev, st);
}

}  // extern "C"
