// Host-buffer entry points of the C ABI: the call a reference-side binding makes with numpy arrays.
// They own their transfers (pageable or pinned host memory both work) and block until the result is
// in the caller's buffers.
#include <stdlib.h>
#include <string.h>

#include <thread>

#include "common.cuh"

namespace slic {

static int d_pad_of(int d) { return (d + 63) / 64 * 64; }

// Pageable source (what clustering/cluster_masks.py:80 hands over: `.cpu().numpy()`): a cudaMemcpyAsync from pageable
// memory is staged by the driver through a small pinned bounce buffer on ONE thread (~10 GB/s measured: 492 MB in ~45 ms,
// twice the level-0 screen it should hide behind).  Here several host threads (3/4 of the cores, at most 12) copy a piece slice by slice into one
// of STAGE_SLOTS pinned staging buffers (kept for the life of the calling thread: cudaHostAlloc costs milliseconds) and
// the DMA of piece c runs while piece c + 1 is being staged.  (Measured at C3: FINCH from a plain numpy array 56 -> 31-32 ms.)
int StagePool::ensure(size_t bytes) {
    if (cap >= bytes) return SLIC_OK;
    for (int i = 0; i < STAGE_SLOTS; ++i) {
        if (busy[i]) SLIC_CUDA_OK(cudaEventSynchronize(done[i]));
        busy[i] = false;
        if (buf[i]) SLIC_CUDA_OK(cudaFreeHost(buf[i]));
        buf[i] = nullptr;
        SLIC_CUDA_OK(cudaHostAlloc(&buf[i], bytes, cudaHostAllocDefault));
        if (!done[i]) SLIC_CUDA_OK(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
    }
    cap = bytes;
    return SLIC_OK;
}
StagePool& stage_pool() {
    static thread_local StagePool pool;
    return pool;
}
bool host_is_pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}
// Page-locked host memory that kernels of the current device can write (cudaHostAlloc / torch pin_memory under unified
// addressing): its device view, else nullptr.  The label matrix of a hierarchy is then written by the stacking kernel
// straight over PCIe - no staging copy in device memory, no second transfer after the level count has reached the host.
void* pinned_device_view(void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
}
static thread_local bool g_stage_threads_shared = false;
void stage_threads_shared() { g_stage_threads_shared = true; }   // this thread is one of several feeding GPUs at the same time
void parallel_host_copy(void* dst, const void* src, size_t bytes) {
    constexpr int STAGE_THREADS_MAX = 32;
    // measured at C3 on a 16-core host (scripts/exp_stage_threads.py): 4 / 8 / 12 / 16 / 24 threads -> 34.9 / 32.3 / 31.2 / 31.4 /
    // 31.9 ms for FINCH(plain numpy array) against 29.8 ms from pinned memory - the host copies 492 MB at ~33 GB/s at best
    // Several GPUs fed from one host at once - the worker threads of slic_finch_multi (stage_threads_shared() below), the
    // ranks of a torchrun job (LOCAL_WORLD_SIZE) - keep 8 threads each, the configuration their numbers were measured with
    // (2 workers x 12 threads on the 16 cores: upload 12.8 -> 14.8 ms).
    static const int default_threads = [] {
        const char* lw = getenv("LOCAL_WORLD_SIZE");
        if (lw && atoi(lw) > 1) return 8;
        const int hw = (int)std::thread::hardware_concurrency();
        const int v = hw > 0 ? hw * 3 / 4 : 8;
        return v < 4 ? 4 : (v > 12 ? 12 : v);
    }();
    int STAGE_THREADS = g_stage_threads_shared ? 8 : default_threads;
    if (const char* e = getenv("SLIC_STAGE_THREADS")) {   // experiments (scripts/exp_stage_threads.py)
        const int v = atoi(e);
        if (v >= 1 && v <= STAGE_THREADS_MAX) STAGE_THREADS = v;
    }
    if (bytes < ((size_t)4 << 20) || STAGE_THREADS == 1) {
        memcpy(dst, src, bytes);
        return;
    }
    std::thread workers[STAGE_THREADS_MAX - 1];
    const size_t per = (bytes / STAGE_THREADS + 4095) / 4096 * 4096;
    for (int t = 1; t < STAGE_THREADS; ++t) {
        const size_t o = per * t, len = o < bytes ? (bytes - o < per ? bytes - o : per) : 0;
        workers[t - 1] = std::thread([=] {
            if (len) memcpy(static_cast<char*>(dst) + o, static_cast<const char*>(src) + o, len);
        });
    }
    memcpy(dst, src, per < bytes ? per : bytes);
    for (int t = 1; t < STAGE_THREADS; ++t) workers[t - 1].join();
}

// host -> device copy of `bytes` on `st`: pinned sources directly, pageable ones through the staging pool in pieces
int copy_to_device_staged(void* dst_dev, const void* src_host, size_t bytes, cudaStream_t st) {
    constexpr size_t PIECE = (size_t)32 << 20;
    if (bytes < ((size_t)8 << 20) || !host_is_pageable(src_host)) {
        SLIC_CUDA_OK(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, st));
        return SLIC_OK;
    }
    StagePool& pool = stage_pool();
    SLIC_PROPAGATE(pool.ensure(PIECE));
    int c = 0;
    for (size_t o = 0; o < bytes; o += PIECE, ++c) {
        const size_t len = bytes - o < PIECE ? bytes - o : PIECE;
        const int slot = c % STAGE_SLOTS;
        if (pool.busy[slot]) SLIC_CUDA_OK(cudaEventSynchronize(pool.done[slot]));
        parallel_host_copy(pool.buf[slot], static_cast<const char*>(src_host) + o, len);
        SLIC_CUDA_OK(cudaMemcpyAsync(static_cast<char*>(dst_dev) + o, pool.buf[slot], len, cudaMemcpyHostToDevice, st));
        SLIC_CUDA_OK(cudaEventRecord(pool.done[slot], st));
        pool.busy[slot] = true;
    }
    return SLIC_OK;
}

}  // namespace slic

extern "C" int slic_copy_to_device(void* dst_dev, const void* src_host, int64_t bytes, slic_stream_t stream) {
    using namespace slic;
    SLIC_REQUIRE(dst_dev && src_host && bytes >= 0, "copy_to_device: bad arguments");
    if (bytes == 0) return SLIC_OK;
    SLIC_PROPAGATE(slic_require_device());
    return copy_to_device_staged(dst_dev, src_host, (size_t)bytes, as_stream(stream));
}

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
        if (e != cudaSuccess) {
            set_error("first_neighbors_host: %s", cudaGetErrorString(e));
            status = SLIC_ERR_CUDA;
        }
        if (status == SLIC_OK) status = copy_to_device_staged(x.ptr, x_host, (size_t)n * d * esz, st);
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
