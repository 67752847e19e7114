// Library identity, error reporting and device queries for the den_b200 C ABI.
#include <stdarg.h>
#include <string.h>

#include "den_common.cuh"

namespace den {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t err, const char* what) {
    set_error("%s: CUDA error %d (%s)", what, (int)err, cudaGetErrorString(err));
    return DEN_ERR_CUDA;
}

int sm_count() {
    static thread_local int cached_dev = -1;
    static thread_local int cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return kSmCountDefault;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = kSmCountDefault;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

}  // namespace den

extern "C" {

int den_version(void) { return DEN_ABI_VERSION; }

const char* den_last_error(void) { return den::g_error; }

int den_device_sm_count(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return den::cuda_fail(e, "den_device_sm_count");
    int n = 0;
    e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return den::cuda_fail(e, "den_device_sm_count");
    return n;
}

}  // extern "C"
