// noflats.cu — K2: the epsilon-sloped ("no flats") depression fill in float64.
//
// Replaces fill.fill_terrain_no_flats (malstroem/algorithms/fill.py:174-232; sweep
// speedups/_fill.pyx:72-124): the fixed point of
//      W(c) = max( z(c), min( min4diag W + diag, min4edge W + short ) ),   W = z on the raster border,
// reached by the reference from W = +inf with raster-wide Gauss-Seidel sweeps.  For short, diag > 0 the
// fixed point is unique, so any schedule that ends in a state where the equation holds
// at every interior cell has the reference's bits.
//
// Schedule used here ("seed, relax tiles from above, certify"):
//   k_nf_init    a dry cell (plain fill F == z) that has a strictly lower filled neighbour is seeded with
//                W = z; every other interior cell starts at +inf (these are the lake / flat cells, 20-35 %
//                of a fractal DEM).  Tiles holding a non-seed cell become active.
//   k_nf_relax   one CTA per active 64x64 tile: tile + 1-cell apron in shared memory, in-place sweeps
//                (alternating column-wise and row-wise ownership, down/up resp. right/left) until the tile
//                is quiet; values only ever decrease.  If the tile's outer ring changed, the neighbouring
//                tiles are activated for the next round.  Rounds repeat until no tile is active.
//   k_nf_verify  one stencil pass checks the equation everywhere.  Relaxed cells satisfy it by
//                construction; a seed can only fail by being too LOW (its lower neighbour is closer than
//                the accumulated epsilons).  Failing seeds are banned and the solve restarts — the result
//                that passes is certified by uniqueness.
#include <math.h>

#include "common.cuh"

namespace ms {

constexpr int NF_T = 64;             // tile edge
constexpr int NF_LD = NF_T + 3;      // shared row stride in doubles (odd: conflict-free column walks)
constexpr int NF_SMEM = (NF_T + 2) * NF_LD * 8 + NF_T * NF_T * 4;

__device__ inline double dmin2(double a, double b) { return a <= b ? a : b; }

__global__ void __launch_bounds__(256) k_nf_init(const float *__restrict__ z, const float *__restrict__ F,
                                                 double *__restrict__ W, const uint8_t *__restrict__ banned,
                                                 uint8_t *tileflag, int rows, int cols, int tiles_x) {
    int c = blockIdx.x * 64 + (threadIdx.x & 63);
    int r = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (r >= rows || c >= cols) return;
    int i = r * cols + c;
    float zc = z[i];
    if (r == 0 || c == 0 || r == rows - 1 || c == cols - 1) {
        W[i] = (double)zc;
        return;
    }
    float f = F[i];
    float m = INFINITY;
#pragma unroll
    for (int dr = -1; dr <= 1; dr++)
#pragma unroll
        for (int dc = -1; dc <= 1; dc++) {
            if (dr == 0 && dc == 0) continue;
            m = fminf(m, __ldg(F + i + dr * cols + dc));
        }
    bool seed = (f == zc) && (m < f) && !(banned && banned[i]);
    W[i] = seed ? (double)zc : (double)INFINITY;
    if (!seed) tileflag[(r / NF_T) * tiles_x + (c / NF_T)] = 1;
}

__global__ void __launch_bounds__(256) k_nf_compact(uint8_t *tileflag, int *list, int *count, int ntiles) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    if (tileflag[t]) {
        tileflag[t] = 0;
        list[atomicAdd(count, 1)] = t;
    }
}

__device__ inline bool nf_update(double *sw, const float *sz, int lr, int lc, double sh, double dg,
                                 int *ring) {
    double *p = sw + (lr + 1) * NF_LD + (lc + 1);
    double w = *p;
    double zc = (double)sz[lr * NF_T + lc];
    if (!(w > zc)) return false;
    double d = dmin2(p[-NF_LD - 1], dmin2(p[-NF_LD + 1], dmin2(p[NF_LD - 1], p[NF_LD + 1])));
    double e = dmin2(p[-NF_LD], dmin2(p[-1], dmin2(p[1], p[NF_LD])));
    double m = dmin2(__dadd_rn(d, dg), __dadd_rn(e, sh));
    m = dmin2(m, w);
    double nv = m >= zc ? m : zc;
    if (nv != w) {
        *p = nv;
        if (lr == 0) ring[0] = 1;
        if (lr == NF_T - 1) ring[1] = 1;
        if (lc == 0) ring[2] = 1;
        if (lc == NF_T - 1) ring[3] = 1;
        return true;
    }
    return false;
}

__global__ void __launch_bounds__(256) k_nf_relax(const float *__restrict__ z, double *W, const int *__restrict__ list,
                                                  uint8_t *tileflag, int rows, int cols, int tiles_x, int tiles_y,
                                                  double sh, double dg) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sw = reinterpret_cast<double *>(smem_raw);
    float *sz = reinterpret_cast<float *>(smem_raw + (NF_T + 2) * NF_LD * 8);
    __shared__ int ring[4];
    int t = list[blockIdx.x];
    int ty = t / tiles_x, tx = t - ty * tiles_x;
    int r0 = ty * NF_T, c0 = tx * NF_T;
    int tid = threadIdx.x;
    if (tid < 4) ring[tid] = 0;
    for (int k = tid; k < (NF_T + 2) * (NF_T + 2); k += 256) {
        int lr = k / (NF_T + 2), lc = k - lr * (NF_T + 2);
        int r = r0 + lr - 1, c = c0 + lc - 1;
        double v = INFINITY;
        if (r >= 0 && r < rows && c >= 0 && c < cols) v = W[(size_t)r * cols + c];
        sw[lr * NF_LD + lc] = v;
    }
    for (int k = tid; k < NF_T * NF_T; k += 256) {
        int lr = k >> 6, lc = k & 63;
        int r = r0 + lr, c = c0 + lc;
        sz[k] = (r < rows && c < cols) ? z[(size_t)r * cols + c] : INFINITY;
    }
    __syncthreads();
    bool any = false;
    for (int iter = 0;; iter++) {
        int changed = 0;
        if ((iter & 1) == 0) {
            // column ownership: thread walks 16 rows of one column down, then up
            int lc = tid & 63, rb = (tid >> 6) * 16;
            for (int k = 0; k < 16; k++) changed |= nf_update(sw, sz, rb + k, lc, sh, dg, ring);
            for (int k = 14; k >= 0; k--) changed |= nf_update(sw, sz, rb + k, lc, sh, dg, ring);
        } else {
            // row ownership: thread walks 16 columns of one row right, then left
            int lr = tid & 63, cb = (tid >> 6) * 16;
            for (int k = 0; k < 16; k++) changed |= nf_update(sw, sz, lr, cb + k, sh, dg, ring);
            for (int k = 14; k >= 0; k--) changed |= nf_update(sw, sz, lr, cb + k, sh, dg, ring);
        }
        // a pass that wrote nothing saw one static state, so every cell is stable: tile converged
        if (!__syncthreads_or(changed)) break;
        any = true;
    }
    if (!any) return;
    for (int k = tid; k < NF_T * NF_T; k += 256) {
        int lr = k >> 6, lc = k & 63;
        int r = r0 + lr, c = c0 + lc;
        if (r < rows && c < cols) W[(size_t)r * cols + c] = sw[(lr + 1) * NF_LD + lc + 1];
    }
    if (tid == 0) {
        bool top = ring[0], bot = ring[1], lef = ring[2], rig = ring[3];
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
                if (!dy && !dx) continue;
                bool hit = (dy < 0 && top) || (dy > 0 && bot) || (dx < 0 && lef) || (dx > 0 && rig);
                // a diagonal neighbour only sees our corner cell: it is covered by either adjoining side
                if (!hit) continue;
                int y = ty + dy, x = tx + dx;
                if (y < 0 || y >= tiles_y || x < 0 || x >= tiles_x) continue;
                tileflag[y * tiles_x + x] = 1;
            }
    }
}

__global__ void __launch_bounds__(256) k_nf_verify(const float *__restrict__ z, const double *__restrict__ W,
                                                   uint8_t *banned, int *nviol, int rows, int cols, double sh,
                                                   double dg) {
    int c = blockIdx.x * 64 + (threadIdx.x & 63);
    int r = blockIdx.y * 4 + (threadIdx.x >> 6);
    int bad = 0;
    if (r > 0 && c > 0 && r < rows - 1 && c < cols - 1) {
        size_t i = (size_t)r * cols + c;
        const double *p = W + i;
        double d = dmin2(p[-cols - 1], dmin2(p[-cols + 1], dmin2(p[cols - 1], p[cols + 1])));
        double e = dmin2(p[-cols], dmin2(p[-1], dmin2(p[1], p[cols])));
        double m = dmin2(__dadd_rn(d, dg), __dadd_rn(e, sh));
        double zc = (double)z[i];
        double g = m >= zc ? m : zc;
        if (g != *p) {
            bad = 1;
            if (banned) banned[i] = 1;
        }
    }
    int cnt = __syncthreads_count(bad);
    if (threadIdx.x == 0 && cnt) atomicAdd(nviol, cnt);
}

int fill_no_flats_dev_impl(const float *dtm, const float *filled, double sh, double dg, double *out,
                           int64_t rows, int64_t cols, int64_t *stats, cudaStream_t s) {
    if (!dtm || !out) { set_error("fill_terrain_no_flats: null pointer"); return MS_ERR_ARG; }
    if (rows < 3 || cols < 3 || rows * cols > (1ll << 30)) {
        set_error("fill_terrain_no_flats: unsupported shape %lld x %lld", (long long)rows, (long long)cols);
        return MS_ERR_SHAPE;
    }
    if (!(sh >= 0) || !(dg >= 0)) { set_error("fill_terrain_no_flats: short/diag must be >= 0"); return MS_ERR_ARG; }
    int64_t n = rows * cols;
    DevBuf<float> ftmp;
    if (!filled) {
        MS_TRY(ftmp.alloc((size_t)n, s));
        MS_TRY(fill_terrain_dev_impl(dtm, ftmp.p, nullptr, rows, cols, nullptr, s));
        filled = ftmp.p;
    }
    static bool attr_done = false;
    if (!attr_done) {
        MS_CUDA(cudaFuncSetAttribute(k_nf_relax, cudaFuncAttributeMaxDynamicSharedMemorySize, NF_SMEM));
        attr_done = true;
    }
    int tiles_x = (int)cdiv(cols, NF_T), tiles_y = (int)cdiv(rows, NF_T);
    int ntiles = tiles_x * tiles_y;
    DevBuf<uint8_t> tileflag, banned;
    DevBuf<int> list, count;
    MS_TRY(tileflag.alloc((size_t)ntiles, s));
    MS_TRY(list.alloc((size_t)ntiles, s));
    MS_TRY(count.alloc(2, s));
    dim3 g2(cdiv(cols, 64), cdiv(rows, 4));
    int64_t *h = host_flags().h;
    int64_t rounds = 0, visits = 0, tries = 0;
    for (;;) {
        tries++;
        MS_CUDA(cudaMemsetAsync(tileflag.p, 0, (size_t)ntiles, s));
        MS_LAUNCH(k_nf_init, g2, 256, 0, s, dtm, filled, out, banned.p, tileflag.p, (int)rows, (int)cols, tiles_x);
        for (;;) {
            MS_CUDA(cudaMemsetAsync(count.p, 0, sizeof(int), s));
            MS_LAUNCH(k_nf_compact, cdiv(ntiles, 256), 256, 0, s, tileflag.p, list.p, count.p, ntiles);
            MS_CUDA(cudaMemcpyAsync(h, count.p, sizeof(int), cudaMemcpyDeviceToHost, s));
            MS_TRY(ms::stream_sync(s));
            int na = *(int *)h;
            if (na == 0) break;
            rounds++;
            visits += na;
            prof_units((int64_t)na * NF_T * NF_T);
            MS_LAUNCH(k_nf_relax, na, 256, NF_SMEM, s, dtm, out, list.p, tileflag.p, (int)rows, (int)cols, tiles_x,
                      tiles_y, sh, dg);
            if (rounds > 4ll * (tiles_x + tiles_y) * NF_T * NF_T) {
                set_error("fill_terrain_no_flats: relaxation did not converge");
                return MS_ERR_NOCONV;
            }
        }
        MS_CUDA(cudaMemsetAsync(count.p + 1, 0, sizeof(int), s));
        MS_LAUNCH(k_nf_verify, g2, 256, 0, s, dtm, out, banned.p, count.p + 1, (int)rows, (int)cols, sh, dg);
        MS_CUDA(cudaMemcpyAsync(h, count.p + 1, sizeof(int), cudaMemcpyDeviceToHost, s));
        MS_TRY(ms::stream_sync(s));
        int nviol = *(int *)h;
        if (nviol == 0) break;
        if (!banned.p) {
            // first failure: allocate the ban map and mark the failing seeds
            MS_TRY(banned.alloc((size_t)n, s));
            MS_CUDA(cudaMemsetAsync(banned.p, 0, (size_t)n, s));
            MS_CUDA(cudaMemsetAsync(count.p + 1, 0, sizeof(int), s));
            MS_LAUNCH(k_nf_verify, g2, 256, 0, s, dtm, out, banned.p, count.p + 1, (int)rows, (int)cols, sh, dg);
        }
        if (tries > 1000) {
            set_error("fill_terrain_no_flats: seed verification did not settle");
            return MS_ERR_NOCONV;
        }
    }
    if (stats) { stats[0] = rounds; stats[1] = visits; stats[2] = tries - 1; }
    return MS_OK;
}

}  // namespace ms

extern "C" {

int ms_fill_terrain_no_flats_dev(const float *dtm, const float *filled, double short_eps, double diag_eps,
                                 double *out, int64_t rows, int64_t cols, int64_t *stats, void *stream) {
    MS_TRY(ms::ensure_init());
    return ms::fill_no_flats_dev_impl(dtm, filled, short_eps, diag_eps, out, rows, cols, stats,
                                      (cudaStream_t)stream);
}

int ms_fill_terrain_no_flats(const float *dtm, double short_eps, double diag_eps, double *out, int64_t rows,
                             int64_t cols) {
    MS_TRY(ms::ensure_init());
    if (!dtm || !out) { ms::set_error("fill_terrain_no_flats: null pointer"); return MS_ERR_ARG; }
    if (rows < 3 || cols < 3 || rows * cols > (1ll << 30)) {
        ms::set_error("fill_terrain_no_flats: unsupported shape %lld x %lld", (long long)rows, (long long)cols);
        return MS_ERR_SHAPE;
    }
    cudaStream_t s = nullptr;
    size_t n = (size_t)(rows * cols);
    ms::DevBuf<float> d;
    ms::DevBuf<double> o;
    MS_TRY(d.alloc(n, s));
    MS_TRY(o.alloc(n, s));
    MS_CUDA(cudaMemcpyAsync(d.p, dtm, n * sizeof(float), cudaMemcpyHostToDevice, s));
    MS_TRY(ms::fill_no_flats_dev_impl(d.p, nullptr, short_eps, diag_eps, o.p, rows, cols, nullptr, s));
    MS_CUDA(cudaMemcpyAsync(out, o.p, n * sizeof(double), cudaMemcpyDeviceToHost, s));
    MS_TRY(ms::stream_sync(s));
    return MS_OK;
}

}  // extern "C"
