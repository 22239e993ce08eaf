// core.cu — library state: device selection, stream-ordered memory pool, error string, pinned flags.
#include <stdarg.h>
#include <string.h>

#include <chrono>
#include <map>
#include <string>
#include <vector>

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

double g_host_t[4] = {0, 0, 0, 0};
double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
int stream_sync(cudaStream_t s) {
    double t0 = now_s();
    cudaError_t e = cudaStreamSynchronize(s);
    g_host_t[0] += now_s() - t0;
    g_host_t[1] += 1;
    if (e != cudaSuccess) {
        set_error("cudaStreamSynchronize -> %s", cudaGetErrorString(e));
        return MS_ERR_CUDA;
    }
    return MS_OK;
}

// ---- scratch arena ---------------------------------------------------------------------------------
struct ArenaBlock { size_t off, size; bool live; };
static char *g_arena = nullptr;
static size_t g_arena_cap = 0, g_arena_top = 0, g_arena_need = 0, g_overflow_live = 0;
static std::vector<ArenaBlock> g_blocks;
static cudaStream_t g_arena_stream = nullptr;
static std::map<void *, size_t> g_overflow;

void *arena_alloc(size_t bytes, cudaStream_t s) {
    bytes = (bytes + 511) & ~(size_t)511;
    if (s != g_arena_stream) {
        // blocks are recycled in stream order: switching streams must wait for the previous one
        if (g_arena) cudaStreamSynchronize(g_arena_stream);
        g_arena_stream = s;
    }
    if (g_blocks.empty() && g_overflow.empty() && g_arena_need > g_arena_cap) {
        if (g_arena) cudaFree(g_arena);
        g_arena = nullptr;
        size_t want = g_arena_need + (g_arena_need >> 3);
        if (cudaMalloc((void **)&g_arena, want) == cudaSuccess) g_arena_cap = want;
        else { g_arena_cap = 0; cudaGetLastError(); }
        g_arena_top = 0;
    }
    size_t virt = g_arena_top + g_overflow_live + bytes;
    if (virt > g_arena_need) g_arena_need = virt;
    if (g_arena && g_arena_top + bytes <= g_arena_cap) {
        g_blocks.push_back({g_arena_top, bytes, true});
        void *p = g_arena + g_arena_top;
        g_arena_top += bytes;
        return p;
    }
    void *p = nullptr;
    if (cudaMallocAsync(&p, bytes, s) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    g_overflow[p] = bytes;
    g_overflow_live += bytes;
    return p;
}

void arena_free(void *p, cudaStream_t s) {
    auto it = g_overflow.find(p);
    if (it != g_overflow.end()) {
        g_overflow_live -= it->second;
        g_overflow.erase(it);
        cudaFreeAsync(p, s);
        return;
    }
    size_t off = (size_t)((char *)p - g_arena);
    for (size_t k = g_blocks.size(); k-- > 0;)
        if (g_blocks[k].off == off) { g_blocks[k].live = false; break; }
    while (!g_blocks.empty() && !g_blocks.back().live) {
        g_arena_top = g_blocks.back().off;
        g_blocks.pop_back();
    }
}

// ---- per-kernel event timing -------------------------------------------------------------------
int g_prof = 0;
int64_t g_prof_units = 0;
struct ProfRec { const char *name; cudaEvent_t a, b; int64_t units; };
static std::vector<ProfRec> g_recs;
static std::vector<cudaEvent_t> g_evpool;

static cudaEvent_t ev_get() {
    if (!g_evpool.empty()) { cudaEvent_t e = g_evpool.back(); g_evpool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

void prof_begin(const char *name, cudaStream_t s) {
    ProfRec r;
    r.name = name; r.a = ev_get(); r.b = ev_get(); r.units = g_prof_units;
    g_prof_units = 0;
    cudaEventRecord(r.a, s);
    g_recs.push_back(r);
}

void prof_end(cudaStream_t s) { cudaEventRecord(g_recs.back().b, s); }

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
    if (ms::g_blocks.empty() && ms::g_arena) {
        cudaFree(ms::g_arena);
        ms::g_arena = nullptr;
        ms::g_arena_cap = ms::g_arena_top = 0;
    }
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

int ms_host_counters(double *out4, int reset) {
    for (int k = 0; k < 4; k++) {
        if (out4) out4[k] = ms::g_host_t[k];
        if (reset) ms::g_host_t[k] = 0;
    }
    return MS_OK;
}

int ms_profile(int enable) {
    // enable > 1: also pre-create that many event pairs now, so that no cudaEventCreate falls inside a timed region
    for (int64_t k = (int64_t)ms::g_evpool.size(); k < 2ll * enable && enable > 1; k++) {
        cudaEvent_t e;
        MS_CUDA(cudaEventCreate(&e));
        ms::g_evpool.push_back(e);
    }
    ms::g_prof = enable ? 1 : 0;
    return MS_OK;
}

// one line per kernel: "<name> <launches> <total ms> <total units>"; clears the records
int ms_profile_report(char *buf, int64_t cap) {
    if (!buf || cap < 1) return MS_ERR_ARG;
    MS_CUDA(cudaDeviceSynchronize());
    struct Agg { int64_t n = 0; double ms = 0; int64_t units = 0; };
    std::map<std::string, Agg> agg;
    for (auto &r : ms::g_recs) {
        float t = 0;
        cudaEventElapsedTime(&t, r.a, r.b);
        Agg &a = agg[r.name];
        a.n++; a.ms += t; a.units += r.units;
        ms::g_evpool.push_back(r.a);
        ms::g_evpool.push_back(r.b);
    }
    ms::g_recs.clear();
    std::string out;
    char line[256];
    for (auto &kv : agg) {
        snprintf(line, sizeof(line), "%s %lld %.6f %lld\n", kv.first.c_str(), (long long)kv.second.n, kv.second.ms,
                 (long long)kv.second.units);
        out += line;
    }
    if ((int64_t)out.size() + 1 > cap) return MS_ERR_ARG;
    memcpy(buf, out.c_str(), out.size() + 1);
    return MS_OK;
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
