// Library-level entry points of the C ABI: error state, version, device check.
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace slic {

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    // a wait inside a gated screen launch timed out (the kernel trapped): say which one
    const char* rec = timeout_record_text();
    if (rec[0]) strncat(g_error, rec, sizeof(g_error) - strlen(g_error) - 1);
}

int num_sms() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, sms = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
            cached = sms;
        else
            return 148;
    }
    return cached;
}

}  // namespace slic

extern "C" {

int slic_abi_version(void) { return SLIC_ABI_VERSION; }

int64_t slic_launch_count(void) { return (int64_t)slic::g_launches.load(std::memory_order_relaxed); }

const char* slic_last_error(void) { return slic::g_error; }

int slic_require_device(void) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        slic::set_error("no CUDA device visible (%s): slic_b200 has no CPU fallback",
                        e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        cudaGetLastError();
        return SLIC_ERR_NO_DEVICE;
    }
    int dev = 0, major = 0;
    SLIC_CUDA_OK(cudaGetDevice(&dev));
    SLIC_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) {
        slic::set_error("device %d has compute capability %d.x; the kernels are built for sm_100a only", dev, major);
        return SLIC_ERR_NO_DEVICE;
    }
    // keep the stream-ordered pool's memory across synchronisations: the default threshold (0) hands every
    // temporary back to the OS at each sync, which costs milliseconds per FINCH level
    static bool pool_ready[64] = {false};
    if (dev < 64 && !pool_ready[dev]) {
        cudaMemPool_t pool;
        SLIC_CUDA_OK(cudaDeviceGetDefaultMemPool(&pool, dev));
        uint64_t keep = UINT64_MAX;
        SLIC_CUDA_OK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        pool_ready[dev] = true;
    }
    return SLIC_OK;
}

}  // extern "C"
