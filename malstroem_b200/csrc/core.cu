// core.cu — library state: device selection, stream-ordered memory pool, error string, pinned flags.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace ms {

static thread_local char g_err[1024] = "";
int64_t g_launches = 0;
static int g_device = -1;
static HostFlags g_flags;

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int HostFlags::init() {
    if (h) return MS_OK;
    MS_CUDA(cudaHostAlloc((void **)&h, 64 * sizeof(int64_t), cudaHostAllocDefault));
    memset(h, 0, 64 * sizeof(int64_t));
    return MS_OK;
}
HostFlags &host_flags() { return g_flags; }

static int init_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error("no CUDA device available (%s): malstroem_b200 has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return MS_ERR_CUDA;
    }
    if (device < 0 || device >= n) {
        set_error("device %d out of range (have %d)", device, n);
        return MS_ERR_ARG;
    }
    MS_CUDA(cudaSetDevice(device));
    // keep freed scratch memory in the pool instead of returning it to the driver after every call
    cudaMemPool_t pool;
    MS_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t thresh = UINT64_MAX;
    MS_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
    MS_TRY(g_flags.init());
    g_device = device;
    return MS_OK;
}

int ensure_init() {
    if (g_device >= 0) {
        // another library (torch) may have switched the current device of this thread
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != g_device) MS_CUDA(cudaSetDevice(g_device));
        return MS_OK;
    }
    int cur = 0;
    if (cudaGetDevice(&cur) != cudaSuccess) cur = 0;
    return init_device(cur);
}

}  // namespace ms

extern "C" {

int ms_version(void) { return 100; }

int ms_init(int device) {
    if (ms::g_device == device) return MS_OK;
    return ms::init_device(device);
}

int ms_shutdown(void) {
    if (ms::g_device < 0) return MS_OK;
    cudaDeviceSynchronize();
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, ms::g_device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
    return MS_OK;
}

const char *ms_last_error(void) { return ms::g_err; }

int ms_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int64_t ms_kernel_launches(int reset) {
    int64_t v = ms::g_launches;
    if (reset) ms::g_launches = 0;
    return v;
}

void *ms_host_alloc(int64_t bytes) {
    void *p = nullptr;
    if (ms::ensure_init() != MS_OK) return nullptr;
    if (cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault) != cudaSuccess) {
        ms::set_error("cudaHostAlloc(%lld) failed", (long long)bytes);
        return nullptr;
    }
    return p;
}

int ms_host_free(void *p) {
    if (p) MS_CUDA(cudaFreeHost(p));
    return MS_OK;
}

}  // extern "C"
