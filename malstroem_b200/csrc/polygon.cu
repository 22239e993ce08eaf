// polygon.cu — SURVEY.md §8(f4): label polygonisation.
//
// Replaces vector.py:42-87 (`vectorize_labels_file`): `gdal.Polygonize(band, band.GetMaskBand(), layer, 0,
// ['8CONNECTED=8'])` — one polygon (an exterior ring and its holes) per 8-connected region of equal cell value; cells
// equal to the band's nodata value (if it has one) belong to no polygon.  GDAL walks the raster two rows at a time on
// one core and stitches edge strings; here every step is a data-parallel pass over cells or over boundary edges:
//
//   1. k_poly_count / scan      per cell: which of its 4 sides face another value (or the raster's edge) -> the cell's
//                               first edge number.  A boundary edge is DIRECTED: a cell's sides are walked clockwise on
//                               the screen (top: east, right: south, bottom: west, left: north), so the cell's region
//                               lies on the right-hand side of every edge.
//   2. k_poly_edges             the successor of every edge, decided at its end vertex from the two cells ahead:
//                               8-connected: ahead-left cell in the region -> turn left (the ring crosses over to the
//                               diagonal neighbour), else ahead-right in the region -> straight on, else turn right
//                               around the cell's corner; 4-connected: the mirror rule.  Every edge has exactly one
//                               successor and one predecessor: the edges fall into disjoint cycles, the rings.
//   3. k_poly_minjump           pointer doubling with a running minimum: every edge learns the lowest edge number of
//                               its ring (the ring's leader).  Synchronous rounds on a ping-pong pair of (leader, jump)
//                               words; a round that changes no leader is the last (proof in DESIGN.md §4 f4).
//   4. k_poly_rankjump          the same doubling with distances: steps from every edge to its ring's leader = its
//                               position in the ring.
//   5. scans + k_poly_place     rings numbered by leader, edges placed at (ring start + position); an edge whose
//                               predecessor runs in another direction starts at a corner: the corners, scanned, are the
//                               ring's vertices in order (k_poly_emit).
//   6. k_poly_union / k_poly_ring   which region a ring belongs to: union-find over the cells (equal value, 4 / 8
//                               neighbours), roots = the region's first cell in raster order.  A ring is an exterior
//                               ring iff its leader is a TOP side (the first cell of the region in raster order), a hole
//                               iff it is a BOTTOM side (the cells above the hole's first row) — no area sums needed.
// The result (ring table + vertex rows / columns as lattice corners) stays on the device until ms_polygonize_fetch
// copies it out; the host side (malstroem_b200/vector.py) groups rings into features and applies the geotransform.
#include <string.h>

#include "common.cuh"

namespace ms {

struct PolyHold {
    int64_t nrings = 0, nverts = 0, nedges = 0;
    int64_t *voff = nullptr;      // nrings + 1
    int32_t *value = nullptr;     // nrings
    int64_t *cell = nullptr;      // nrings: the leader's cell
    int64_t *region = nullptr;    // nrings: first cell of the ring's region
    uint8_t *hole = nullptr;      // nrings
    int32_t *vrow = nullptr, *vcol = nullptr;      // nverts
};
static PolyHold g_poly;

void poly_release() {
    cudaFree(g_poly.voff); cudaFree(g_poly.value); cudaFree(g_poly.cell); cudaFree(g_poly.region);
    cudaFree(g_poly.hole); cudaFree(g_poly.vrow); cudaFree(g_poly.vcol);
    g_poly = PolyHold();
}

struct PolyGrid {
    const int32_t *lab;
    int rows, cols;
    int connect8, has_nodata;
    int32_t nodata;
};

__device__ inline bool poly_same(const PolyGrid &g, int r, int c, int32_t v) {
    return r >= 0 && r < g.rows && c >= 0 && c < g.cols && g.lab[(size_t)r * g.cols + c] == v;
}

// sides of the cell that are boundary edges: bit 0 top, 1 right, 2 bottom, 3 left
__device__ inline int poly_mask(const PolyGrid &g, int r, int c, int32_t v) {
    if (g.has_nodata && v == g.nodata) return 0;
    return (poly_same(g, r - 1, c, v) ? 0 : 1) | (poly_same(g, r, c + 1, v) ? 0 : 2) | (poly_same(g, r + 1, c, v) ? 0 : 4) |
           (poly_same(g, r, c - 1, v) ? 0 : 8);
}

__global__ void __launch_bounds__(256) k_poly_count(PolyGrid g, int *cnt) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)g.rows * g.cols) return;
    int r = (int)(i / g.cols), c = (int)(i - (int64_t)r * g.cols);
    cnt[i] = __popc(poly_mask(g, r, c, g.lab[i]));
}

// heading of side s: 0 east, 1 south, 2 west, 3 north; forward and right-hand vectors as (dr, dc)
__device__ __constant__ const int kFR[4] = {0, 1, 0, -1}, kFC[4] = {1, 0, -1, 0};
__device__ __constant__ const int kRR[4] = {1, 0, -1, 0}, kRC[4] = {0, -1, 0, 1};

__global__ void __launch_bounds__(256) k_poly_edges(PolyGrid g, const int *__restrict__ eoff, uint32_t *ecell,
                                                    int *succ, unsigned long long *state) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)g.rows * g.cols) return;
    int r = (int)(i / g.cols), c = (int)(i - (int64_t)r * g.cols);
    const int32_t v = g.lab[i];
    const int mask = poly_mask(g, r, c, v);
    if (!mask) return;
    int e = eoff[i];
    for (int s = 0; s < 4; s++) {
        if (!(mask & (1 << s))) continue;
        // the cells ahead of the end vertex: ahead-right continues on the region's side, ahead-left on the other
        const int arr = r + kFR[s], arc = c + kFC[s];
        const int alr = arr - kRR[s], alc = arc - kRC[s];
        const bool in_ar = poly_same(g, arr, arc, v), in_al = poly_same(g, alr, alc, v);
        int nr, nc, ns;
        bool left, straight;
        if (g.connect8) { left = in_al; straight = !left && in_ar; }
        else { left = in_ar && in_al; straight = in_ar && !in_al; }
        if (left) { nr = alr; nc = alc; ns = (s + 3) & 3; }
        else if (straight) { nr = arr; nc = arc; ns = s; }
        else { nr = r; nc = c; ns = (s + 1) & 3; }
        const int64_t j = (int64_t)nr * g.cols + nc;
        const int nmask = (nr == r && nc == c) ? mask : poly_mask(g, nr, nc, v);
        const int q = eoff[j] + __popc(nmask & ((1 << ns) - 1));
        ecell[e] = (uint32_t)((i << 2) | s);
        succ[e] = q;
        state[e] = ((unsigned long long)(unsigned)e << 32) | (unsigned)q;      // (leader so far, jump)
        e++;
    }
}

// hi word: lowest edge number among the 2^k edges from here on; lo word: the edge 2^k steps ahead
__global__ void __launch_bounds__(256) k_poly_minjump(const unsigned long long *__restrict__ a, unsigned long long *b,
                                                      int n, int *changed) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const unsigned long long x = a[e];
    const unsigned long long y = a[(unsigned)x];
    const unsigned lx = (unsigned)(x >> 32), ly = (unsigned)(y >> 32);
    const unsigned l = lx < ly ? lx : ly;
    b[e] = ((unsigned long long)l << 32) | (unsigned)y;
    if (l != lx) *changed = 1;
}

// hi word: steps from here to the jump target; lo word: the jump target (the leader points at itself, 0 steps)
__global__ void __launch_bounds__(256) k_poly_rankinit(const unsigned long long *__restrict__ lead_state,
                                                       const int *__restrict__ succ, unsigned long long *c, int n) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const unsigned lead = (unsigned)(lead_state[e] >> 32);
    c[e] = lead == (unsigned)e ? (unsigned long long)(unsigned)e : ((1ull << 32) | (unsigned)succ[e]);
}

__global__ void __launch_bounds__(256) k_poly_rankjump(const unsigned long long *__restrict__ a, unsigned long long *b,
                                                       int n, int *changed) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const unsigned long long x = a[e];
    const unsigned long long y = a[(unsigned)x];
    b[e] = (((x >> 32) + (y >> 32)) << 32) | (unsigned)y;
    if ((unsigned)y != (unsigned)x) *changed = 1;
}

// leaders: flag + ring length (scanned into ring number / ring start)
__global__ void __launch_bounds__(256) k_poly_leaders(const unsigned long long *__restrict__ lead_state,
                                                      const unsigned long long *__restrict__ rank,
                                                      const int *__restrict__ succ, int *isleader, int *ringlen, int n) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const bool ld = (unsigned)(lead_state[e] >> 32) == (unsigned)e;
    isleader[e] = ld;
    ringlen[e] = ld ? (int)(rank[succ[e]] >> 32) + 1 : 0;
}

// every edge goes to (start of its ring + position in the ring); it tells its successor's slot whether that one
// starts at a corner
__global__ void __launch_bounds__(256) k_poly_place(const unsigned long long *__restrict__ lead_state,
                                                    const unsigned long long *__restrict__ rank,
                                                    const int *__restrict__ succ, const uint32_t *__restrict__ ecell,
                                                    const int *__restrict__ rstart, const int *__restrict__ ringlen,
                                                    uint32_t *placed, int *corner, int n) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const unsigned lead = (unsigned)(lead_state[e] >> 32);
    const int L = ringlen[lead], base = rstart[lead];
    const int d = (int)(rank[e] >> 32);
    const int pos = d == 0 ? 0 : L - d;
    placed[base + pos] = ecell[e];
    const int q = succ[e];
    corner[base + (pos + 1 == L ? 0 : pos + 1)] = (ecell[q] & 3u) != (ecell[e] & 3u);
}

__global__ void __launch_bounds__(256) k_poly_emit(const uint32_t *__restrict__ placed, const int *__restrict__ corner,
                                                   const int *__restrict__ vidx, int cols, int32_t *vrow, int32_t *vcol,
                                                   int n) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n || !corner[k]) return;
    const uint32_t ec = placed[k];
    const int s = (int)(ec & 3u);
    const int64_t i = (int64_t)(ec >> 2);
    const int r = (int)(i / cols), c = (int)(i - (int64_t)r * cols);
    // start vertex of the side: top (r, c), right (r, c + 1), bottom (r + 1, c + 1), left (r + 1, c)
    const int v = vidx[k];
    vrow[v] = r + (s >= 2 ? 1 : 0);
    vcol[v] = c + ((s == 1 || s == 2) ? 1 : 0);
}

// ---- regions: union-find over the cells, roots = lowest cell index (every write is an atomicMin towards an ancestor)
__device__ inline int poly_root(int *par, int x) {
    for (;;) {
        int p = *(volatile int *)(par + x);
        if (p == x) return x;
        int gp = *(volatile int *)(par + p);
        if (gp != p) atomicMin(par + x, gp);
        x = p;
    }
}

__device__ inline void poly_unite(int *par, int a, int b) {
    for (;;) {
        a = poly_root(par, a);
        b = poly_root(par, b);
        if (a == b) return;
        const int hi = a > b ? a : b, lo = a > b ? b : a;
        const int old = atomicMin(par + hi, lo);
        if (old == hi) return;
        a = old;
        b = lo;
    }
}

__global__ void __launch_bounds__(256) k_poly_parinit(int *par, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) par[i] = (int)i;
}

__global__ void __launch_bounds__(256) k_poly_union(PolyGrid g, int *par) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)g.rows * g.cols) return;
    int r = (int)(i / g.cols), c = (int)(i - (int64_t)r * g.cols);
    const int32_t v = g.lab[i];
    if (g.has_nodata && v == g.nodata) return;
    if (poly_same(g, r, c - 1, v)) poly_unite(par, (int)i, (int)i - 1);
    if (poly_same(g, r - 1, c, v)) poly_unite(par, (int)i, (int)i - g.cols);
    if (g.connect8) {
        // a diagonal neighbour only matters when neither cell between the two joins them already
        if (poly_same(g, r - 1, c - 1, v) && !poly_same(g, r - 1, c, v) && !poly_same(g, r, c - 1, v))
            poly_unite(par, (int)i, (int)i - g.cols - 1);
        if (poly_same(g, r - 1, c + 1, v) && !poly_same(g, r - 1, c, v)) poly_unite(par, (int)i, (int)i - g.cols + 1);
    }
}

__global__ void __launch_bounds__(256) k_poly_ring(PolyGrid g, const int *__restrict__ isleader,
                                                   const int *__restrict__ ridx, const int *__restrict__ rstart,
                                                   const int *__restrict__ vidx, const uint32_t *__restrict__ ecell,
                                                   int *par, int64_t *voff, int32_t *value, int64_t *cell, int64_t *region,
                                                   uint8_t *hole, int n, int64_t nrings, int64_t nverts) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e == 0) voff[nrings] = nverts;
    if (e >= n || !isleader[e]) return;
    const int k = ridx[e];
    const uint32_t ec = ecell[e];
    const int64_t i = (int64_t)(ec >> 2);
    voff[k] = vidx[rstart[e]];
    value[k] = g.lab[i];
    cell[k] = i;
    region[k] = poly_root(par, (int)i);
    hole[k] = (ec & 3u) != 0u;        // exterior rings start with a top side, holes with a bottom side
}

// synchronous doubling rounds on a ping-pong pair; *a holds the result
static int jump_rounds(bool rank, unsigned long long **a, unsigned long long **b, int n, cudaStream_t s) {
    DevBuf<int> flag;
    MS_TRY(flag.alloc(1, s));
    int64_t *h = host_flags().h;
    for (int round = 0;; round++) {
        if (round > 40) { set_error("polygonize: ring %s did not converge", rank ? "positions" : "leaders"); return MS_ERR_NOCONV; }
        MS_CUDA(cudaMemsetAsync(flag.p, 0, sizeof(int), s));
        if (rank) MS_LAUNCH(k_poly_rankjump, cdiv(n, 256), 256, 0, s, (const unsigned long long *)*a, *b, n, flag.p);
        else MS_LAUNCH(k_poly_minjump, cdiv(n, 256), 256, 0, s, (const unsigned long long *)*a, *b, n, flag.p);
        MS_TRY(ms::readback(h, flag.p, sizeof(int), s));
        MS_TRY(stream_sync(s));
        unsigned long long *t = *a; *a = *b; *b = t;
        if (*(int *)h == 0) return MS_OK;
    }
}

template <class T>
static int hold_alloc(T **p, size_t count) {
    if (cudaMalloc((void **)p, (count ? count : 1) * sizeof(T)) != cudaSuccess) {
        cudaGetLastError();
        set_error("polygonize: cudaMalloc of %zu bytes failed", count * sizeof(T));
        return MS_ERR_CUDA;
    }
    return MS_OK;
}

int polygonize_dev_impl(const int32_t *labels, int64_t rows, int64_t cols, int connect8, int has_nodata, int32_t nodata,
                        int64_t *counts3, cudaStream_t s) {
    if (!labels || !counts3 || rows < 1 || cols < 1 || rows * cols > (1ll << 30)) {
        set_error("polygonize: unsupported shape %lld x %lld", (long long)rows, (long long)cols);
        return MS_ERR_ARG;
    }
    poly_release();
    const int64_t n = rows * cols;
    PolyGrid g{labels, (int)rows, (int)cols, connect8 ? 1 : 0, has_nodata ? 1 : 0, nodata};
    int64_t *h = host_flags().h;
    DevBuf<int> eoff;
    DevBuf<int64_t> total;
    MS_TRY(eoff.alloc((size_t)n, s));
    MS_TRY(total.alloc(1, s));
    MS_LAUNCH(k_poly_count, cdiv(n, 256), 256, 0, s, g, eoff.p);
    {
        // the scan's block sums are int32: the edge count has to fit
        MS_TRY(exclusive_scan_i32(eoff.p, eoff.p, n, total.p, s));
        MS_TRY(ms::readback(h, total.p, sizeof(int64_t), s));
        MS_TRY(stream_sync(s));
    }
    const int64_t E = h[0];
    if (E < 0 || E >= (1ll << 31) - 1024) { set_error("polygonize: %lld boundary edges do not fit", (long long)E); return MS_ERR_ARG; }
    counts3[0] = counts3[1] = 0;
    counts3[2] = E;
    if (E == 0) {        // everything is nodata
        MS_TRY(hold_alloc(&g_poly.voff, 1));
        MS_CUDA(cudaMemsetAsync(g_poly.voff, 0, sizeof(int64_t), s));
        MS_TRY(stream_sync(s));
        return MS_OK;
    }
    const int ne = (int)E;
    DevBuf<uint32_t> ecell, placed;
    DevBuf<int> succ, isleader, ringlen, par;
    DevBuf<unsigned long long> sa, sb, ra, rb;
    MS_TRY(ecell.alloc((size_t)ne, s));
    MS_TRY(succ.alloc((size_t)ne, s));
    MS_TRY(sa.alloc((size_t)ne, s));
    MS_TRY(sb.alloc((size_t)ne, s));
    MS_LAUNCH(k_poly_edges, cdiv(n, 256), 256, 0, s, g, (const int *)eoff.p, ecell.p, succ.p, sa.p);
    unsigned long long *a = sa.p, *b = sb.p;
    MS_TRY(jump_rounds(false, &a, &b, ne, s));
    unsigned long long *lead_state = a;       // hi word = the ring's leader; the other buffer is free again
    MS_TRY(ra.alloc((size_t)ne, s));
    MS_LAUNCH(k_poly_rankinit, cdiv(ne, 256), 256, 0, s, (const unsigned long long *)lead_state, (const int *)succ.p, ra.p, ne);
    unsigned long long *c = ra.p, *d = b;
    MS_TRY(jump_rounds(true, &c, &d, ne, s));
    unsigned long long *rank = c;
    // the regions (independent of the rings; the ring table needs their roots)
    MS_TRY(par.alloc((size_t)n, s));
    MS_LAUNCH(k_poly_parinit, cdiv(n, 256), 256, 0, s, par.p, n);
    MS_LAUNCH(k_poly_union, cdiv(n, 256), 256, 0, s, g, par.p);
    // ring numbers and ring starts
    MS_TRY(isleader.alloc((size_t)ne, s));
    MS_TRY(ringlen.alloc((size_t)ne, s));
    MS_LAUNCH(k_poly_leaders, cdiv(ne, 256), 256, 0, s, (const unsigned long long *)lead_state, (const unsigned long long *)rank,
              (const int *)succ.p, isleader.p, ringlen.p, ne);
    DevBuf<int> ridx, rstart, corner, vidx;
    MS_TRY(ridx.alloc((size_t)ne, s));
    MS_TRY(rstart.alloc((size_t)ne, s));
    MS_TRY(exclusive_scan_i32(isleader.p, ridx.p, ne, total.p, s));
    MS_TRY(ms::readback(h, total.p, sizeof(int64_t), s));
    MS_TRY(stream_sync(s));
    const int64_t nrings = h[0];
    MS_TRY(exclusive_scan_i32(ringlen.p, rstart.p, ne, total.p, s));
    MS_TRY(placed.alloc((size_t)ne, s));
    MS_TRY(corner.alloc((size_t)ne, s));
    MS_TRY(vidx.alloc((size_t)ne, s));
    MS_LAUNCH(k_poly_place, cdiv(ne, 256), 256, 0, s, (const unsigned long long *)lead_state, (const unsigned long long *)rank,
              (const int *)succ.p, (const uint32_t *)ecell.p, (const int *)rstart.p, (const int *)ringlen.p, placed.p, corner.p, ne);
    MS_TRY(exclusive_scan_i32(corner.p, vidx.p, ne, total.p, s));
    MS_TRY(ms::readback(h, total.p, sizeof(int64_t), s));
    MS_TRY(stream_sync(s));
    const int64_t nverts = h[0];
    MS_TRY(hold_alloc(&g_poly.voff, (size_t)nrings + 1));
    MS_TRY(hold_alloc(&g_poly.value, (size_t)nrings));
    MS_TRY(hold_alloc(&g_poly.cell, (size_t)nrings));
    MS_TRY(hold_alloc(&g_poly.region, (size_t)nrings));
    MS_TRY(hold_alloc(&g_poly.hole, (size_t)nrings));
    MS_TRY(hold_alloc(&g_poly.vrow, (size_t)nverts));
    MS_TRY(hold_alloc(&g_poly.vcol, (size_t)nverts));
    MS_LAUNCH(k_poly_emit, cdiv(ne, 256), 256, 0, s, (const uint32_t *)placed.p, (const int *)corner.p, (const int *)vidx.p,
              (int)cols, g_poly.vrow, g_poly.vcol, ne);
    MS_LAUNCH(k_poly_ring, cdiv(ne, 256), 256, 0, s, g, (const int *)isleader.p, (const int *)ridx.p, (const int *)rstart.p,
              (const int *)vidx.p, (const uint32_t *)ecell.p, par.p, g_poly.voff, g_poly.value, g_poly.cell, g_poly.region,
              g_poly.hole, ne, nrings, nverts);
    MS_TRY(stream_sync(s));
    g_poly.nrings = nrings;
    g_poly.nverts = nverts;
    g_poly.nedges = E;
    counts3[0] = nrings;
    counts3[1] = nverts;
    return MS_OK;
}

}  // namespace ms

extern "C" {

int ms_polygonize_dev(const int32_t *labels, int64_t rows, int64_t cols, int connect8, int has_nodata, int32_t nodata,
                      int64_t *counts3, void *stream) {
    MS_TRY(ms::ensure_init());
    return ms::polygonize_dev_impl(labels, rows, cols, connect8, has_nodata, nodata, counts3, (cudaStream_t)stream);
}

int ms_polygonize(const int32_t *labels, int64_t rows, int64_t cols, int connect8, int has_nodata, int32_t nodata,
                  int64_t *counts3) {
    MS_TRY(ms::ensure_init());
    if (!labels || !counts3 || rows < 1 || cols < 1) { ms::set_error("polygonize: bad argument"); return MS_ERR_ARG; }
    cudaStream_t s = nullptr;
    ms::HostCall hc;
    int32_t *l = nullptr;
    MS_TRY(hc.in(labels, (size_t)(rows * cols), s, &l));
    return ms::polygonize_dev_impl(l, rows, cols, connect8, has_nodata, nodata, counts3, s);
}

/* copies the rings of the last ms_polygonize* call to host arrays (sizes from its counts) and releases them */
int ms_polygonize_fetch(int64_t *ring_vertex_offset, int32_t *ring_value, int64_t *ring_cell, int64_t *ring_region,
                        uint8_t *ring_hole, int32_t *vertex_row, int32_t *vertex_col) {
    MS_TRY(ms::ensure_init());
    ms::PolyHold &p = ms::g_poly;
    if (!p.voff) { ms::set_error("polygonize_fetch: no result is held (ms_polygonize first)"); return MS_ERR_ARG; }
    if (!ring_vertex_offset || (p.nrings && (!ring_value || !ring_cell || !ring_region || !ring_hole)) ||
        (p.nverts && (!vertex_row || !vertex_col))) {
        ms::set_error("polygonize_fetch: null pointer");
        return MS_ERR_ARG;
    }
    cudaStream_t s = nullptr;
    MS_CUDA(cudaMemcpyAsync(ring_vertex_offset, p.voff, (size_t)(p.nrings + 1) * 8, cudaMemcpyDeviceToHost, s));
    if (p.nrings) {
        MS_CUDA(cudaMemcpyAsync(ring_value, p.value, (size_t)p.nrings * 4, cudaMemcpyDeviceToHost, s));
        MS_CUDA(cudaMemcpyAsync(ring_cell, p.cell, (size_t)p.nrings * 8, cudaMemcpyDeviceToHost, s));
        MS_CUDA(cudaMemcpyAsync(ring_region, p.region, (size_t)p.nrings * 8, cudaMemcpyDeviceToHost, s));
        MS_CUDA(cudaMemcpyAsync(ring_hole, p.hole, (size_t)p.nrings, cudaMemcpyDeviceToHost, s));
    }
    if (p.nverts) {
        MS_CUDA(cudaMemcpyAsync(vertex_row, p.vrow, (size_t)p.nverts * 4, cudaMemcpyDeviceToHost, s));
        MS_CUDA(cudaMemcpyAsync(vertex_col, p.vcol, (size_t)p.nverts * 4, cudaMemcpyDeviceToHost, s));
    }
    MS_TRY(ms::stream_sync(s));
    ms::poly_release();
    return MS_OK;
}

}  // extern "C"
