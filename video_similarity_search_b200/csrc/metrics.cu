// Cluster-quality scores on the device: the mutual information, the two label
// entropies and the expected mutual information that normalized_mutual_info_score / adjusted_mutual_info_score
// need - the two sklearn calls online_train.py:633-642 makes on rank 0 right after every clustering.
//
// Arithmetic follows scikit-learn (third-party, not under /root/reference; requirements.txt:5 pins 0.22.0):
//   sklearn/metrics/cluster/_supervised.py  mutual_info_score (sparse-contingency branch), entropy
//   sklearn/metrics/cluster/_expected_mutual_info_fast.pyx  expected_mutual_information
// All integer work (contingency counts) is exact; the float64 sums are reduced in a fixed order, so results are
// reproducible run to run.  HBM-bound: 8 n bytes of labels + 4 R C bytes of contingency cells, read once.
//
// EMI.  sklearn loops over all R x C (row sum a_i, column sum b_j) pairs and all feasible n_ij - O(R C min(a, b))
// terms, each with three lgamma calls: ~10 s on one core at Kinetics size (R = 400, C = 21 436).  The term depends
// on (a_i, b_j) only through their VALUES, so here the pairs of DISTINCT values are enumerated, weighted by their
// multiplicities, with lgamma(k + 1) tabulated once for k = 0..n; one warp per value pair.
#include <math.h>

#include "common.cuh"
#include "primitives.cuh"

namespace slic {

constexpr int MT_THREADS = 256;

__global__ void __launch_bounds__(MT_THREADS) contingency_count_kernel(const int* __restrict__ lt, const int* __restrict__ lp,
                                                                      int64_t n, int num_true, int num_pred,
                                                                      int* __restrict__ cells, int* __restrict__ a,
                                                                      int* __restrict__ b, int* __restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int t = lt[i], p = lp[i];
    if (t < 0 || t >= num_true || p < 0 || p >= num_pred) {   // labels are documented dense; a direct C caller may not comply
        atomicAdd(bad, 1);
        return;
    }
    atomicAdd(cells + (int64_t)t * num_pred + p, 1);
    atomicAdd(a + t, 1);
    atomicAdd(b + p, 1);
}

// labels outside [0, num) were met: every output becomes NaN (the scores are undefined), [4] carries the count negated
__global__ void poison_scores_kernel(const int* __restrict__ bad, double* __restrict__ out) {
    if (*bad > 0) {
        for (int i = 0; i < 6; ++i) out[i] = nan("");
        out[4] = -(double)*bad;
    }
}

// fixed-order block sum: warp shuffles, then warp 0 over the warp totals
__device__ __forceinline__ double block_sum(double v, double* smem) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (warp == 0) {
        t = lane < (blockDim.x >> 5) ? smem[lane] : 0.0;
        t = warp_sum(t);
    }
    return t;   // valid in warp 0
}

// sklearn entropy(): -sum (p_i / N) (log p_i - log N) over non-empty classes, 0 for a single class.
// out[0] = entropy, out[1] = number of non-empty classes.  One block.
__global__ void __launch_bounds__(MT_THREADS) entropy_kernel(const int* __restrict__ counts, int num, double n_total,
                                                            double* __restrict__ out) {
    __shared__ double sm[MT_THREADS / 32];
    double h = 0.0, nz = 0.0;
    const double log_n = log(n_total);
    for (int i = threadIdx.x; i < num; i += blockDim.x) {
        const int c = counts[i];
        if (c > 0) {
            const double pc = (double)c;
            h += (pc / n_total) * (log(pc) - log_n);
            nz += 1.0;
        }
    }
    h = block_sum(h, sm);
    nz = block_sum(nz, sm);
    if (threadIdx.x == 0) {
        out[1] = nz;
        out[0] = nz <= 1.0 ? 0.0 : -h;
    }
}

// mutual_info_score on the non-zero cells: partial[block] = sum of
//   (n_ij / N) (log n_ij - log N) + (n_ij / N) (-log(a_i b_j) + log N + log N),   terms below float64 eps dropped
__global__ void __launch_bounds__(MT_THREADS) mi_partial_kernel(const int* __restrict__ cells, int64_t num_cells, int num_pred,
                                                               const int* __restrict__ a, const int* __restrict__ b,
                                                               double n_total, double* __restrict__ partial) {
    __shared__ double sm[MT_THREADS / 32];
    const double log_n = log(n_total);
    double acc = 0.0;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < num_cells; c += (int64_t)gridDim.x * blockDim.x) {
        const int nij = cells[c];
        if (nij == 0) continue;
        const int64_t i = c / num_pred, j = c - i * num_pred;
        const double cnm = (double)nij / n_total;
        const double outer = (double)((long long)a[i] * (long long)b[j]);
        const double log_outer = -log(outer) + log_n + log_n;
        double term = cnm * (log((double)nij) - log_n) + cnm * log_outer;
        if (fabs(term) < 2.220446049250313e-16) term = 0.0;
        acc += term;
    }
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

__global__ void __launch_bounds__(MT_THREADS) sum_partials_kernel(const double* __restrict__ partial, int num, double* out,
                                                                 int clip_at_zero) {
    __shared__ double sm[MT_THREADS / 32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < num; i += blockDim.x) acc += partial[i];
    acc = block_sum(acc, sm);
    if (threadIdx.x == 0) *out = (clip_at_zero && acc < 0.0) ? 0.0 : acc;
}

__global__ void lgamma_table_kernel(double* __restrict__ table, int64_t n) {   // table[k] = lgamma(k + 1), k = 0..n
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k <= n) table[k] = lgamma((double)k + 1.0);
}

__global__ void flag_positive_kernel(const int* __restrict__ mult, int64_t bins, int* __restrict__ flags) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v < bins) flags[v] = (v > 0 && mult[v] > 0) ? 1 : 0;
}
__global__ void compact_values_kernel(const int* __restrict__ mult, const int* __restrict__ flags,
                                      const int* __restrict__ pos, int64_t bins, int* __restrict__ vals,
                                      int* __restrict__ mults) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v < bins && flags[v]) {
        vals[pos[v]] = (int)v;
        mults[pos[v]] = mult[v];
    }
}

// expected_mutual_information: one warp per pair of distinct (a, b) values, lanes over the feasible n_ij.
__global__ void __launch_bounds__(MT_THREADS) emi_partial_kernel(const int* __restrict__ va, const int* __restrict__ ma,
                                                                const int* __restrict__ num_a,
                                                                const int* __restrict__ vb, const int* __restrict__ mb,
                                                                const int* __restrict__ num_b,
                                                                const double* __restrict__ T, int64_t n,
                                                                double* __restrict__ partial) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t num_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int ua = *num_a, ub = *num_b;
    const int64_t pairs = (int64_t)ua * ub;
    const double dn = (double)n, log_n = log(dn), t_n = T[n];
    double acc = 0.0;
    for (int64_t p = warp; p < pairs; p += num_warps) {
        const int ia = (int)(p / ub), ib = (int)(p - (int64_t)ia * ub);
        const int64_t a = va[ia], b = vb[ib];
        const double weight = (double)ma[ia] * (double)mb[ib];
        const double log_ab = log((double)a) + log((double)b);
        const double g0 = T[a] + T[b] + T[n - a] + T[n - b] - t_n;
        const int64_t start = a + b - n > 1 ? a + b - n : 1;
        const int64_t end = a < b ? a : b;   // inclusive
        double s = 0.0;
        for (int64_t nij = start + lane; nij <= end; nij += 32) {
            const double term1 = (double)nij / dn;
            const double term2 = log_n + log((double)nij) - log_ab;
            const double gln = g0 - T[nij] - T[a - nij] - T[b - nij] - T[n - a - b + nij];
            s += term1 * term2 * exp(gln);
        }
        acc += weight * s;
    }
    acc = warp_sum(acc);
    if (lane == 0) partial[warp] = acc;
}

}  // namespace slic

extern "C" int slic_cluster_metrics(const int32_t* labels_true_dev, const int32_t* labels_pred_dev, int64_t n,
                                    int32_t num_true, int32_t num_pred, int32_t want_emi, double* out_dev,
                                    slic_stream_t stream) {
    using namespace slic;
    SLIC_REQUIRE(n >= 1 && n < ((int64_t)1 << 31) && num_true >= 1 && num_pred >= 1, "cluster_metrics: bad shape");
    SLIC_REQUIRE(labels_true_dev && labels_pred_dev && out_dev, "cluster_metrics: null pointer");
    const int64_t cells = (int64_t)num_true * num_pred;
    if (cells > ((int64_t)1 << 28)) {
        set_error("cluster_metrics: %d x %d contingency cells exceed the dense limit of 2^28", num_true, num_pred);
        return SLIC_ERR_UNSUPPORTED;
    }
    SLIC_PROPAGATE(slic_require_device());
    cudaStream_t st = as_stream(stream);
    Scratch hist, a, b, partial, bad;
    SLIC_CUDA_OK(bad.alloc(sizeof(int), st));
    SLIC_CUDA_OK(cudaMemsetAsync(bad.ptr, 0, sizeof(int), st));
    SLIC_CUDA_OK(hist.alloc(cells * sizeof(int), st));
    SLIC_CUDA_OK(a.alloc((size_t)num_true * sizeof(int), st));
    SLIC_CUDA_OK(b.alloc((size_t)num_pred * sizeof(int), st));
    SLIC_CUDA_OK(cudaMemsetAsync(hist.ptr, 0, cells * sizeof(int), st));
    SLIC_CUDA_OK(cudaMemsetAsync(a.ptr, 0, (size_t)num_true * sizeof(int), st));
    SLIC_CUDA_OK(cudaMemsetAsync(b.ptr, 0, (size_t)num_pred * sizeof(int), st));
    SLIC_CUDA_OK(cudaMemsetAsync(out_dev, 0, 6 * sizeof(double), st));
    contingency_count_kernel<<<(unsigned)ceil_div(n, MT_THREADS), MT_THREADS, 0, st>>>(labels_true_dev, labels_pred_dev, n,
                                                                                      num_true, num_pred, hist.as<int>(),
                                                                                      a.as<int>(), b.as<int>(), bad.as<int>());
    SLIC_LAUNCH_OK();
    // out: [0] mi  [1] h_true  [2] h_pred  [3] emi  [4] non-empty true classes  [5] non-empty predicted clusters
    Scratch ent;
    SLIC_CUDA_OK(ent.alloc(4 * sizeof(double), st));
    entropy_kernel<<<1, MT_THREADS, 0, st>>>(a.as<int>(), num_true, (double)n, ent.as<double>());
    SLIC_LAUNCH_OK();
    entropy_kernel<<<1, MT_THREADS, 0, st>>>(b.as<int>(), num_pred, (double)n, ent.as<double>() + 2);
    SLIC_LAUNCH_OK();
    SLIC_CUDA_OK(cudaMemcpyAsync(out_dev + 1, ent.as<double>() + 0, sizeof(double), cudaMemcpyDeviceToDevice, st));
    SLIC_CUDA_OK(cudaMemcpyAsync(out_dev + 2, ent.as<double>() + 2, sizeof(double), cudaMemcpyDeviceToDevice, st));
    SLIC_CUDA_OK(cudaMemcpyAsync(out_dev + 4, ent.as<double>() + 1, sizeof(double), cudaMemcpyDeviceToDevice, st));
    SLIC_CUDA_OK(cudaMemcpyAsync(out_dev + 5, ent.as<double>() + 3, sizeof(double), cudaMemcpyDeviceToDevice, st));
    int64_t blocks = ceil_div(cells, (int64_t)MT_THREADS * 8);
    const int64_t max_blocks = (int64_t)num_sms() * 8;
    if (blocks > max_blocks) blocks = max_blocks;
    const int64_t emi_blocks = (int64_t)num_sms() * 4, emi_warps = emi_blocks * (MT_THREADS / 32);
    SLIC_CUDA_OK(partial.alloc((size_t)(blocks > emi_warps ? blocks : emi_warps) * sizeof(double), st));
    mi_partial_kernel<<<(unsigned)blocks, MT_THREADS, 0, st>>>(hist.as<int>(), cells, num_pred, a.as<int>(), b.as<int>(),
                                                               (double)n, partial.as<double>());
    SLIC_LAUNCH_OK();
    sum_partials_kernel<<<1, MT_THREADS, 0, st>>>(partial.as<double>(), (int)blocks, out_dev + 0, 1);   // np.clip(mi.sum(), 0, None)
    SLIC_LAUNCH_OK();
    if (want_emi) {
        const int64_t bins = n + 1;
        Scratch table, mult, flags, pos, va, ma, vb, mb, totals;
        SLIC_CUDA_OK(table.alloc((size_t)bins * sizeof(double), st));
        SLIC_CUDA_OK(mult.alloc((size_t)bins * sizeof(int), st));
        SLIC_CUDA_OK(flags.alloc((size_t)bins * sizeof(int), st));
        SLIC_CUDA_OK(pos.alloc((size_t)bins * sizeof(int), st));
        SLIC_CUDA_OK(va.alloc((size_t)num_true * sizeof(int), st));
        SLIC_CUDA_OK(ma.alloc((size_t)num_true * sizeof(int), st));
        SLIC_CUDA_OK(vb.alloc((size_t)num_pred * sizeof(int), st));
        SLIC_CUDA_OK(mb.alloc((size_t)num_pred * sizeof(int), st));
        SLIC_CUDA_OK(totals.alloc(2 * sizeof(int), st));
        lgamma_table_kernel<<<(unsigned)ceil_div(bins, MT_THREADS), MT_THREADS, 0, st>>>(table.as<double>(), n);
        SLIC_LAUNCH_OK();
        for (int side = 0; side < 2; ++side) {
            const int* sums = side == 0 ? a.as<int>() : b.as<int>();
            const int64_t num = side == 0 ? num_true : num_pred;
            // multiplicity of every row-sum (column-sum) value, then the list of distinct positive values
            SLIC_PROPAGATE(histogram_i32(sums, num, mult.as<int>(), bins, st));
            flag_positive_kernel<<<(unsigned)ceil_div(bins, MT_THREADS), MT_THREADS, 0, st>>>(mult.as<int>(), bins,
                                                                                             flags.as<int>());
            SLIC_LAUNCH_OK();
            SLIC_PROPAGATE(exclusive_scan_i32(flags.as<int>(), pos.as<int>(), bins, totals.as<int>() + side, st));
            compact_values_kernel<<<(unsigned)ceil_div(bins, MT_THREADS), MT_THREADS, 0, st>>>(
                mult.as<int>(), flags.as<int>(), pos.as<int>(), bins, side == 0 ? va.as<int>() : vb.as<int>(),
                side == 0 ? ma.as<int>() : mb.as<int>());
            SLIC_LAUNCH_OK();
        }
        emi_partial_kernel<<<(unsigned)emi_blocks, MT_THREADS, 0, st>>>(va.as<int>(), ma.as<int>(), totals.as<int>(),
                                                                        vb.as<int>(), mb.as<int>(), totals.as<int>() + 1,
                                                                        table.as<double>(), n, partial.as<double>());
        SLIC_LAUNCH_OK();
        sum_partials_kernel<<<1, MT_THREADS, 0, st>>>(partial.as<double>(), (int)emi_warps, out_dev + 3, 0);
        SLIC_LAUNCH_OK();
    }
    poison_scores_kernel<<<1, 1, 0, st>>>(bad.as<int>(), out_dev);
    SLIC_LAUNCH_OK();
    return SLIC_OK;
}
