// Host-buffer entry points of the C ABI: the call a reference-side binding makes with numpy arrays.
// They own their transfers (pageable or pinned host memory both work) and block until the result is
// in the caller's buffers.
#include "common.cuh"

namespace slic {

static int d_pad_of(int d) { return (d + 63) / 64 * 64; }

}  // namespace slic

extern "C" int slic_first_neighbors_host(const void* x_host, int64_t n, int32_t d, int32_t dtype, int32_t* nn_out_host,
                                         void* dist_out_host) {
    using namespace slic;
    SLIC_REQUIRE(x_host && nn_out_host && n > 1 && d > 0, "first_neighbors_host: bad arguments");
    SLIC_REQUIRE(dtype == SLIC_F32 || dtype == SLIC_F64, "first_neighbors_host: bad dtype");
    SLIC_PROPAGATE(slic_require_device());
    const size_t esz = dtype == SLIC_F32 ? 4 : 8;
    const int dp = d_pad_of(d);
    cudaStream_t st = nullptr;
    SLIC_CUDA_OK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    int status = SLIC_OK;
    {
        Scratch x, unit, ub, nn, dist;
        cudaError_t e = x.alloc((size_t)n * d * esz, st);
        if (e == cudaSuccess) e = unit.alloc((size_t)n * d * esz, st);
        if (e == cudaSuccess) e = ub.alloc((size_t)n * dp * 2, st);
        if (e == cudaSuccess) e = nn.alloc((size_t)n * 4, st);
        if (e == cudaSuccess) e = dist.alloc((size_t)n * esz, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(x.ptr, x_host, (size_t)n * d * esz, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) {
            set_error("first_neighbors_host: %s", cudaGetErrorString(e));
            status = SLIC_ERR_CUDA;
        }
        if (status == SLIC_OK)
            status = slic_normalize_rows(x.ptr, n, d, dtype, unit.ptr, nullptr, ub.as<uint16_t>(), dp, st);
        if (status == SLIC_OK) {
            if (n >= 2048)
                status = slic_nn_top1(unit.ptr, ub.as<uint16_t>(), n, unit.ptr, ub.as<uint16_t>(), n, d, dp, dtype, 0,
                                      0.f, nn.as<int32_t>(), dist.ptr, nullptr, st);
            else
                status = slic_nn_exact_top1(unit.ptr, nullptr, n, unit.ptr, n, d, dtype, 0, nn.as<int32_t>(), dist.ptr,
                                            st);
        }
        if (status == SLIC_OK) {
            e = cudaMemcpyAsync(nn_out_host, nn.ptr, (size_t)n * 4, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess && dist_out_host)
                e = cudaMemcpyAsync(dist_out_host, dist.ptr, (size_t)n * esz, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) {
                set_error("first_neighbors_host: %s", cudaGetErrorString(e));
                status = SLIC_ERR_CUDA;
            }
        }
    }
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    return status;
}
