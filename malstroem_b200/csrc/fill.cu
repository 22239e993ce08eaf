// fill.cu — K0 (min/max) and K1: plain depression fill as a catchment graph + Boruvka contraction.
//
// Replaces fill.fill_terrain (malstroem/algorithms/fill.py:112-171; sweeps speedups/_fill.pyx:28-70),
// whose result is the greatest fixed point of W = max(z, min over the 8 neighbours of W) with W = z on
// the raster border, i.e. W(c) = the lowest "highest cell" over all paths from c to the border.  Only
// min/max of input values are involved, so any algorithm that reaches the fixed point is bit-exact.
//
//   1. k_descent      every interior cell points at the smallest cell of its 3x3 window in the strict
//                     total order (z, flat index); border cells point at cell 0 ("outside")
//   2. forest_resolve pointer jumping -> every cell knows the local minimum ("catchment") it drains to
//   3. scan           dense catchment ids (0 = outside)
//   4. Boruvka rounds on the catchment graph (edge weight = max(z_a, z_b) over adjacent cells of two
//                     catchments): every component not yet merged with "outside" finds its lowest
//                     outgoing edge (one raster pass, 64-bit atomicMin of (weight, edge id)), hooks to
//                     the other side, and every catchment in it raises E = max(E, that weight).
//                     Claim: when the component reaches "outside", E is the catchment's
//                     spill elevation.
//   5. k_fill_final   filled = max(z, E[catchment]), depths = filled - z
#include "common.cuh"

namespace ms {

constexpr unsigned long long KEY_NONE = ~0ull;

// ---- K0 -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_minmax(const float *z, int64_t n, uint32_t *keys) {
    uint32_t lo = 0xffffffffu, hi = 0u;
    int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
    for (; i < n; i += stride) {
        if (i + 4 <= n) {
            float4 v = *reinterpret_cast<const float4 *>(z + i);
            uint32_t a = okey32(v.x), b = okey32(v.y), c = okey32(v.z), d = okey32(v.w);
            lo = min(lo, min(min(a, b), min(c, d)));
            hi = max(hi, max(max(a, b), max(c, d)));
        } else {
            for (int64_t j = i; j < n; j++) {
                uint32_t a = okey32(z[j]);
                lo = min(lo, a);
                hi = max(hi, a);
            }
        }
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    __shared__ uint32_t slo[8], shi[8];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { slo[w] = lo; shi[w] = hi; }
    __syncthreads();
    if (w == 0) {
        lo = lane < 8 ? slo[lane] : 0xffffffffu;
        hi = lane < 8 ? shi[lane] : 0u;
        lo = __reduce_min_sync(0xffffffffu, lo);
        hi = __reduce_max_sync(0xffffffffu, hi);
        if (lane == 0) {
            atomicMin(&keys[0], lo);
            atomicMax(&keys[1], hi);
        }
    }
}

__global__ void k_minmax_finish(const uint32_t *keys, float *out) {
    out[0] = okey32_inv(keys[0]);
    out[1] = okey32_inv(keys[1]);
}

int minmax_dev(const float *z, int64_t n, float *out2, cudaStream_t s) {
    DevBuf<uint32_t> keys;
    MS_TRY(keys.alloc(2, s));
    static const uint32_t init[2] = {0xffffffffu, 0u};
    MS_CUDA(cudaMemcpyAsync(keys.p, init, sizeof(init), cudaMemcpyHostToDevice, s));
    int64_t want = (n / 4 + 255) / 256;
    int blocks = (int)(want < 1 ? 1 : (want > 148 * 16 ? 148 * 16 : want));
    MS_LAUNCH(k_minmax, blocks, 256, 0, s, z, n, keys.p);
    MS_LAUNCH(k_minmax_finish, 1, 1, 0, s, keys.p, out2);
    return MS_OK;
}

// ---- K1 -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_descent(const float *__restrict__ z, int *__restrict__ ptr, int rows,
                                                 int cols) {
    int c = blockIdx.x * 64 + (threadIdx.x & 63);
    int r = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (r >= rows || c >= cols) return;
    int i = r * cols + c;
    if (r == 0 || c == 0 || r == rows - 1 || c == cols - 1) {
        ptr[i] = 0;
        return;
    }
    float bz = z[i];
    int bi = i;
#pragma unroll
    for (int dr = -1; dr <= 1; dr++)
#pragma unroll
        for (int dc = -1; dc <= 1; dc++) {
            if (dr == 0 && dc == 0) continue;
            int j = i + dr * cols + dc;
            float zj = __ldg(z + j);
            if (zj < bz || (zj == bz && j < bi)) { bz = zj; bi = j; }
        }
    ptr[i] = bi;
}

__global__ void __launch_bounds__(256) k_rootflag(const int *ptr, int *flag, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (ptr[i] == (int)i) ? 1 : 0;
}

// lab[i] = cid[ptr[i]], written over ptr (each thread only overwrites its own slot; cid is separate)
__global__ void __launch_bounds__(256) k_catchment_ids(int *ptr_lab, const int *cid, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ptr_lab[i] = cid[ptr_lab[i]];
}

__global__ void __launch_bounds__(256) k_boruvka_init(int *comp, uint32_t *E, int nC) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l < nC) {
        comp[l] = l;
        E[l] = okey32(-INFINITY);
    }
}

// lowest outgoing edge of every component that has not reached "outside" (component 0) yet
__global__ void __launch_bounds__(256) k_minedge(const float *__restrict__ z, const int *__restrict__ lab,
                                                 const int *__restrict__ comp, unsigned long long *best,
                                                 int rows, int cols) {
    int c = blockIdx.x * 64 + (threadIdx.x & 63);
    int r = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (r >= rows || c >= cols) return;
    int i = r * cols + c;
    int l = lab[i];
    if (l == 0) return;
    int cc = comp[l];
    if (cc == 0) return;
    float zc = z[i];
    unsigned long long bk = KEY_NONE;
    // interior cell (label != 0 implies not on the border): all 8 neighbours are in the raster
#pragma unroll
    for (int dr = -1; dr <= 1; dr++)
#pragma unroll
        for (int dc = -1; dc <= 1; dc++) {
            if (dr == 0 && dc == 0) continue;
            int j = i + dr * cols + dc;
            int lj = __ldg(lab + j);
            if (lj == l) continue;
            if (__ldg(comp + lj) == cc) continue;
            float w = fmaxf(zc, __ldg(z + j));
            // symmetric edge id: lower cell index and the direction to the higher one (E, SW, S, SE)
            int lo = j < i ? j : i;
            int code = (dr == 0) ? 0 : ((dr * dc == -1) ? 1 : (dc == 0 ? 2 : 3));
            unsigned long long key = ((unsigned long long)okey32(w) << 32) | (unsigned)(((unsigned)lo << 2) | code);
            bk = key < bk ? key : bk;
        }
    if (bk != KEY_NONE && bk < best[cc]) atomicMin(&best[cc], bk);
}

__device__ inline int edge_other(int lo, int code, int cols) {
    return lo + (code == 0 ? 1 : (code == 1 ? cols - 1 : (code == 2 ? cols : cols + 1)));
}

// every live component hooks to the component across its lowest edge; mutual pairs keep the smaller id
__global__ void __launch_bounds__(256) k_hook(const unsigned long long *__restrict__ best,
                                              const int *__restrict__ comp, const int *__restrict__ lab,
                                              int *parent, uint32_t *wk, int nC, int cols) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nC) return;
    parent[k] = k;
    wk[k] = okey32(-INFINITY);
    if (k == 0 || comp[k] != k) return;
    unsigned long long b = best[k];
    if (b == KEY_NONE) return;
    unsigned id = (unsigned)(b & 0xffffffffu);
    int lo = (int)(id >> 2), code = (int)(id & 3u);
    int hi = edge_other(lo, code, cols);
    int ca = comp[lab[lo]], cb = comp[lab[hi]];
    int t = (ca == k) ? cb : ca;
    wk[k] = (uint32_t)(b >> 32);
    if (t != 0 && best[t] == b && k < t) t = k;
    parent[k] = t;
}

__device__ inline int uf_find(int *parent, int k) {
    int p = parent[k];
    for (;;) {
        int g = parent[p];
        if (g == p) return p;
        parent[k] = g;
        k = p;
        p = g;
    }
}

// E = max(E, weight of the edge the catchment's component left through); component := merged root
__global__ void __launch_bounds__(256) k_boruvka_update(int *comp, uint32_t *E, int *parent,
                                                        const uint32_t *__restrict__ wk, int nC,
                                                        unsigned long long *best, int *remaining) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    int live = 0;
    if (l < nC) {
        best[l] = KEY_NONE;
        int k = comp[l];
        if (k != 0) {
            uint32_t e = E[l], w = wk[k];
            if (w > e) E[l] = w;
            int r = uf_find(parent, k);
            comp[l] = r;
            live = (r == l);      // still the representative of a component that has not reached 0
        }
    }
    int cnt = __syncthreads_count(live);
    if (threadIdx.x == 0 && cnt) atomicAdd(remaining, cnt);
}

__global__ void __launch_bounds__(256) k_fill_final(const float *__restrict__ z, const int *__restrict__ lab,
                                                    const uint32_t *__restrict__ E, float *__restrict__ filled,
                                                    float *__restrict__ depths, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float zc = z[i];
    float w = fmaxf(zc, okey32_inv(__ldg(E + lab[i])));
    filled[i] = w;
    if (depths) depths[i] = __fsub_rn(w, zc);
}

int fill_terrain_dev_impl(const float *dtm, float *filled, float *depths, int64_t rows, int64_t cols,
                          int64_t *stats, cudaStream_t s) {
    if (!dtm || !filled) { set_error("fill_terrain: null pointer"); return MS_ERR_ARG; }
    if (rows < 3 || cols < 3 || rows * cols > (1ll << 30)) {
        set_error("fill_terrain: unsupported shape %lld x %lld", (long long)rows, (long long)cols);
        return MS_ERR_SHAPE;
    }
    int64_t n = rows * cols;
    DevBuf<int> lab, tmp;
    DevBuf<int64_t> total;
    MS_TRY(lab.alloc((size_t)n, s));
    MS_TRY(tmp.alloc((size_t)n, s));
    MS_TRY(total.alloc(1, s));
    dim3 g2(cdiv(cols, 64), cdiv(rows, 4));
    unsigned g1 = cdiv(n, 256);
    int64_t *h = host_flags().h;

    MS_LAUNCH(k_descent, g2, 256, 0, s, dtm, lab.p, (int)rows, (int)cols);
    int64_t jump_rounds = 0;
    MS_TRY(forest_resolve(lab.p, n, &jump_rounds, s));
    MS_LAUNCH(k_rootflag, g1, 256, 0, s, lab.p, tmp.p, n);
    MS_TRY(exclusive_scan_i32(tmp.p, tmp.p, n, total.p, s));
    MS_LAUNCH(k_catchment_ids, g1, 256, 0, s, lab.p, tmp.p, n);
    MS_CUDA(cudaMemcpyAsync(h, total.p, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    MS_TRY(ms::stream_sync(s));
    int nC = (int)h[0];
    tmp.release();

    DevBuf<int> comp, parent, remaining;
    DevBuf<uint32_t> E, wk;
    DevBuf<unsigned long long> best;
    MS_TRY(comp.alloc(nC, s));
    MS_TRY(parent.alloc(nC, s));
    MS_TRY(E.alloc(nC, s));
    MS_TRY(wk.alloc(nC, s));
    MS_TRY(best.alloc(nC, s));
    MS_TRY(remaining.alloc(1, s));
    unsigned gc = cdiv(nC, 256);
    MS_LAUNCH(k_boruvka_init, gc, 256, 0, s, comp.p, E.p, nC);
    MS_CUDA(cudaMemsetAsync(best.p, 0xff, (size_t)nC * sizeof(unsigned long long), s));
    int rounds = 0;
    int64_t live = nC - 1;       // live components; each round every one of them merges with another
    while (live > 0) {
        MS_LAUNCH(k_minedge, g2, 256, 0, s, dtm, lab.p, comp.p, best.p, (int)rows, (int)cols);
        MS_LAUNCH(k_hook, gc, 256, 0, s, best.p, comp.p, lab.p, parent.p, wk.p, nC, (int)cols);
        MS_CUDA(cudaMemsetAsync(remaining.p, 0, sizeof(int), s));
        MS_LAUNCH(k_boruvka_update, gc, 256, 0, s, comp.p, E.p, parent.p, wk.p, nC, best.p, remaining.p);
        MS_CUDA(cudaMemcpyAsync(h, remaining.p, sizeof(int), cudaMemcpyDeviceToHost, s));
        MS_TRY(ms::stream_sync(s));
        int64_t now = *(int *)h;
        rounds++;
        if (now >= live || rounds > 64) {
            set_error("fill_terrain: Boruvka contraction stalled (%lld -> %lld live components, round %d)",
                      (long long)live, (long long)now, rounds);
            return MS_ERR_NOCONV;
        }
        live = now;
    }
    MS_LAUNCH(k_fill_final, g1, 256, 0, s, dtm, lab.p, E.p, filled, depths, n);
    if (stats) { stats[0] = rounds; stats[1] = nC; stats[5] = jump_rounds; }
    return MS_OK;
}

}  // namespace ms

extern "C" {

int ms_minmax_f32_dev(const float *dem, int64_t n, float *out_minmax_dev2, void *stream) {
    MS_TRY(ms::ensure_init());
    if (!dem || !out_minmax_dev2 || n <= 0) { ms::set_error("minmax: bad argument"); return MS_ERR_ARG; }
    return ms::minmax_dev(dem, n, out_minmax_dev2, (cudaStream_t)stream);
}

int ms_minmax_f32(const float *dem, int64_t n, float *out_min, float *out_max) {
    MS_TRY(ms::ensure_init());
    if (!dem || n <= 0) { ms::set_error("minmax: bad argument"); return MS_ERR_ARG; }
    cudaStream_t s = nullptr;
    ms::DevBuf<float> d, o;
    MS_TRY(d.alloc((size_t)n, s));
    MS_TRY(o.alloc(2, s));
    MS_CUDA(cudaMemcpyAsync(d.p, dem, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, s));
    MS_TRY(ms::minmax_dev(d.p, n, o.p, s));
    float r[2];
    MS_CUDA(cudaMemcpyAsync(r, o.p, sizeof(r), cudaMemcpyDeviceToHost, s));
    MS_TRY(ms::stream_sync(s));
    *out_min = r[0];
    *out_max = r[1];
    return MS_OK;
}

int ms_fill_terrain_dev(const float *dtm, float *filled, float *depths, int64_t rows, int64_t cols,
                        void *stream) {
    MS_TRY(ms::ensure_init());
    return ms::fill_terrain_dev_impl(dtm, filled, depths, rows, cols, nullptr, (cudaStream_t)stream);
}

int ms_fill_terrain(const float *dtm, float *filled, float *depths, int64_t rows, int64_t cols) {
    MS_TRY(ms::ensure_init());
    if (!dtm || !filled) { ms::set_error("fill_terrain: null pointer"); return MS_ERR_ARG; }
    if (rows < 3 || cols < 3 || rows * cols > (1ll << 30)) {
        ms::set_error("fill_terrain: unsupported shape %lld x %lld", (long long)rows, (long long)cols);
        return MS_ERR_SHAPE;
    }
    cudaStream_t s = nullptr;
    size_t n = (size_t)(rows * cols);
    ms::DevBuf<float> d, f, dp;
    MS_TRY(d.alloc(n, s));
    MS_TRY(f.alloc(n, s));
    if (depths) MS_TRY(dp.alloc(n, s));
    MS_CUDA(cudaMemcpyAsync(d.p, dtm, n * sizeof(float), cudaMemcpyHostToDevice, s));
    MS_TRY(ms::fill_terrain_dev_impl(d.p, f.p, depths ? dp.p : nullptr, rows, cols, nullptr, s));
    MS_CUDA(cudaMemcpyAsync(filled, f.p, n * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (depths) MS_CUDA(cudaMemcpyAsync(depths, dp.p, n * sizeof(float), cudaMemcpyDeviceToHost, s));
    MS_TRY(ms::stream_sync(s));
    return MS_OK;
}

}  // extern "C"
