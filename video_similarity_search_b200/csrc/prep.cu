// K1 prep: row norms, unit rows in the reference dtype, and the f16 copy for the tensor-core
// screen.  HBM-bound: one pass over X (read n*d*sizeof(T), write n*d*sizeof(T) + n*d_pad*2).
// Follows sklearn.preprocessing.normalize as cosine_similarity applies it behind
// clustering/finch.py:27 / evaluate.py:213 / iic_retrieve_clips.py:295: norm = sqrt(sum x^2),
// zero norms replaced by 1, division in the array dtype.
#include <cuda_fp16.h>

#include "common.cuh"

namespace slic {

// 128-thread CTAs (4 rows): one warp per SM sub-partition, so that these CTAs fit next to the persistent screen kernel
// while it waits for the chunk they are normalising (gated launches, finch_driver.cu)
constexpr int NORM_THREADS = 128;
template <typename T>
__global__ void __launch_bounds__(NORM_THREADS) normalize_rows_kernel(const T* __restrict__ x, int64_t n, int d,
                                                             T* __restrict__ unit, T* __restrict__ norms,
                                                             __half* __restrict__ ub, int d_pad) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const T* xr = x + row * d;
    double acc = 0.0;
    for (int k = lane; k < d; k += 32) {
        double v = (double)xr[k];
        acc = fma(v, v, acc);
    }
    acc = warp_sum(acc);
    // the reference takes the square root in the array dtype
    T nrm = sizeof(T) == 4 ? (T)sqrtf((float)acc) : (T)sqrt(acc);
    if (nrm == (T)0) nrm = (T)1;
    if (norms && lane == 0) norms[row] = nrm;
    for (int k = lane; k < d_pad; k += 32) {
        T u = (T)0;
        if (k < d) {
            u = xr[k] / nrm;
            if (unit) unit[row * d + k] = u;
        }
        if (ub) ub[row * d_pad + k] = __float2half_rn((float)u);
    }
}

// ---- column centring (coclr_classify.py:788-789: feature - feature.mean(dim=0, keepdim=True)) ----------------
// HBM-bound: x is read twice (sums, subtract) and written once.  Column sums are taken in float64 over fixed row slabs
// and the slab partials added in slab order: deterministic.
constexpr int CENTER_SLAB = 256;
__global__ void __launch_bounds__(256) column_partial_sums_kernel(const float* __restrict__ x, int64_t n, int d,
                                                                 double* __restrict__ partial) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= d) return;
    const int64_t r0 = (int64_t)blockIdx.y * CENTER_SLAB;
    const int64_t r1 = r0 + CENTER_SLAB < n ? r0 + CENTER_SLAB : n;
    // consecutive threads read consecutive columns (coalesced); four independent chains keep loads in flight
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int64_t r = r0;
    for (; r + 4 <= r1; r += 4) {
        a0 += (double)x[r * d + col];
        a1 += (double)x[(r + 1) * d + col];
        a2 += (double)x[(r + 2) * d + col];
        a3 += (double)x[(r + 3) * d + col];
    }
    for (; r < r1; ++r) a0 += (double)x[r * d + col];
    partial[(int64_t)blockIdx.y * d + col] = (a0 + a1) + (a2 + a3);
}
// one warp per column: lanes stride the slabs, fixed-order shuffle reduction
__global__ void __launch_bounds__(256) column_means_kernel(const double* __restrict__ partial, int64_t slabs, int d, int64_t n,
                                                          float* __restrict__ means) {
    const int lane = threadIdx.x & 31;
    const int col = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (col >= d) return;
    double acc = 0.0;
    for (int64_t s = lane; s < slabs; s += 32) acc += partial[s * d + col];
    acc = warp_sum(acc);
    if (lane == 0) means[col] = (float)(acc / (double)n);
}
__global__ void __launch_bounds__(256) subtract_columns_kernel(const float* __restrict__ x, int64_t total, int d,
                                                              const float* __restrict__ means, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) out[i] = x[i] - means[i % d];
}
// d % 4 == 0 and 16-byte aligned rows: 128-bit loads and stores
__global__ void __launch_bounds__(256) subtract_columns_v4_kernel(const float4* __restrict__ x, int64_t total4, int d4,
                                                                 const float4* __restrict__ means, float4* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const float4 v = x[i], m = __ldg(means + (i % d4));
    out[i] = make_float4(v.x - m.x, v.y - m.y, v.z - m.z, v.w - m.w);
}

int normalize_kernel_shape(int* threads, int* regs) {
    cudaFuncAttributes fa;
    SLIC_CUDA_OK(cudaFuncGetAttributes(&fa, normalize_rows_kernel<float>));
    *threads = NORM_THREADS;
    *regs = fa.numRegs;
    return SLIC_OK;
}

}  // namespace slic

extern "C" int slic_normalize_rows(const void* x_dev, int64_t n, int32_t d, int32_t dtype, void* unit_dev,
                                   void* norms_dev, uint16_t* unit_f16_dev, int32_t d_pad, slic_stream_t stream) {
    SLIC_REQUIRE(n >= 0 && d > 0, "normalize_rows: bad shape");
    SLIC_REQUIRE(dtype == SLIC_F32 || dtype == SLIC_F64, "normalize_rows: dtype must be SLIC_F32 or SLIC_F64");
    if (unit_f16_dev) SLIC_REQUIRE(d_pad >= d && d_pad % 64 == 0, "normalize_rows: d_pad must be a multiple of 64 >= d");
    if (n == 0) return SLIC_OK;
    const int dp = unit_f16_dev ? d_pad : d;
    const unsigned blocks = (unsigned)slic::ceil_div(n, slic::NORM_THREADS / 32);
    cudaStream_t st = slic::as_stream(stream);
    if (dtype == SLIC_F32)
        slic::normalize_rows_kernel<float><<<blocks, slic::NORM_THREADS, 0, st>>>((const float*)x_dev, n, d, (float*)unit_dev,
                                                                   (float*)norms_dev, (__half*)unit_f16_dev, dp);
    else
        slic::normalize_rows_kernel<double><<<blocks, slic::NORM_THREADS, 0, st>>>((const double*)x_dev, n, d, (double*)unit_dev,
                                                                    (double*)norms_dev, (__half*)unit_f16_dev, dp);
    SLIC_LAUNCH_OK();
    return SLIC_OK;
}

extern "C" int slic_center_columns(const float* x_dev, int64_t n, int32_t d, float* out_dev, float* means_out_dev,
                                   slic_stream_t stream) {
    using namespace slic;
    SLIC_REQUIRE(n >= 1 && d >= 1 && x_dev && out_dev, "center_columns: bad arguments");
    cudaStream_t st = as_stream(stream);
    const int64_t slabs = ceil_div(n, CENTER_SLAB);
    SLIC_REQUIRE(slabs <= 65535, "center_columns: more than 16.7 M rows");
    Scratch partial, means;
    SLIC_CUDA_OK(partial.alloc((size_t)slabs * d * sizeof(double), st));
    SLIC_CUDA_OK(means.alloc((size_t)d * sizeof(float), st));
    dim3 grid((unsigned)ceil_div(d, 256), (unsigned)slabs);
    column_partial_sums_kernel<<<grid, 256, 0, st>>>(x_dev, n, d, partial.as<double>());
    SLIC_LAUNCH_OK();
    column_means_kernel<<<(unsigned)ceil_div((int64_t)d * 32, 256), 256, 0, st>>>(partial.as<double>(), slabs, d, n,
                                                                                  means.as<float>());
    SLIC_LAUNCH_OK();
    const bool v4 = d % 4 == 0 && ((reinterpret_cast<uintptr_t>(x_dev) | reinterpret_cast<uintptr_t>(out_dev)) & 15) == 0;
    if (v4)
        subtract_columns_v4_kernel<<<(unsigned)ceil_div(n * d / 4, 256), 256, 0, st>>>(
            (const float4*)x_dev, n * d / 4, d / 4, (const float4*)means.as<float>(), (float4*)out_dev);
    else
        subtract_columns_kernel<<<(unsigned)ceil_div(n * d, 256), 256, 0, st>>>(x_dev, n * d, d, means.as<float>(), out_dev);
    SLIC_LAUNCH_OK();
    if (means_out_dev) SLIC_CUDA_OK(cudaMemcpyAsync(means_out_dev, means.ptr, (size_t)d * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return SLIC_OK;
}
