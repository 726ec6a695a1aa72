// K1 prep: row norms, unit rows in the reference dtype, and the bf16 copy for the tensor-core
// screen.  HBM-bound: one pass over X (read n*d*sizeof(T), write n*d*sizeof(T) + n*d_pad*2).
// Follows sklearn.preprocessing.normalize as cosine_similarity applies it behind
// clustering/finch.py:27 / evaluate.py:213 / iic_retrieve_clips.py:295: norm = sqrt(sum x^2),
// zero norms replaced by 1, division in the array dtype.
#include <cuda_bf16.h>

#include "common.cuh"

namespace slic {

// 128-thread CTAs (4 rows): one warp per SM sub-partition, so that these CTAs fit next to the persistent screen kernel
// while it waits for the chunk they are normalising (gated launches, finch_driver.cu)
constexpr int NORM_THREADS = 128;
template <typename T>
__global__ void __launch_bounds__(NORM_THREADS) normalize_rows_kernel(const T* __restrict__ x, int64_t n, int d,
                                                             T* __restrict__ unit, T* __restrict__ norms,
                                                             __nv_bfloat16* __restrict__ ub, int d_pad) {
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
        if (ub) ub[row * d_pad + k] = __float2bfloat16_rn((float)u);
    }
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
                                   void* norms_dev, uint16_t* unit_bf16_dev, int32_t d_pad, slic_stream_t stream) {
    SLIC_REQUIRE(n >= 0 && d > 0, "normalize_rows: bad shape");
    SLIC_REQUIRE(dtype == SLIC_F32 || dtype == SLIC_F64, "normalize_rows: dtype must be SLIC_F32 or SLIC_F64");
    if (unit_bf16_dev) SLIC_REQUIRE(d_pad >= d && d_pad % 64 == 0, "normalize_rows: d_pad must be a multiple of 64 >= d");
    if (n == 0) return SLIC_OK;
    const int dp = unit_bf16_dev ? d_pad : d;
    const unsigned blocks = (unsigned)slic::ceil_div(n, slic::NORM_THREADS / 32);
    cudaStream_t st = slic::as_stream(stream);
    if (dtype == SLIC_F32)
        slic::normalize_rows_kernel<float><<<blocks, slic::NORM_THREADS, 0, st>>>((const float*)x_dev, n, d, (float*)unit_dev,
                                                                   (float*)norms_dev, (__nv_bfloat16*)unit_bf16_dev, dp);
    else
        slic::normalize_rows_kernel<double><<<blocks, slic::NORM_THREADS, 0, st>>>((const double*)x_dev, n, d, (double*)unit_dev,
                                                                    (double*)norms_dev, (__nv_bfloat16*)unit_bf16_dev, dp);
    SLIC_LAUNCH_OK();
    return SLIC_OK;
}
