// noflats.cu — K2: the epsilon-sloped ("no flats") depression fill in float64.
//
// Replaces fill.fill_terrain_no_flats (malstroem/algorithms/fill.py:174-232; sweep
// speedups/_fill.pyx:72-124): the fixed point of
//      W(c) = max( z(c), min( min4diag W + diag, min4edge W + short ) ),   W = z on the raster border,
// reached by the reference from W = +inf with raster-wide Gauss-Seidel sweeps.  For short, diag > 0 the
// fixed point is unique, so any schedule that ends in a state where the equation holds
// at every interior cell has the reference's bits.
//
// Schedule used here ("seed, relax from above with a persistent tile solver, certify"):
//   k_nf_init    a dry cell (plain fill F == z) that has a strictly lower filled neighbour is seeded with
//                W = z; every other interior cell starts at +inf (these are the lake / flat cells, 20-35 %
//                of a fractal DEM).  Tiles holding a non-seed cell become active.
//   k_nf_solve   ONE cooperative launch, CTAs resident on every SM.  Rounds: the CTAs take the active 64x64
//                tiles off a list (ticket counter); a tile + 1-cell apron is held in shared memory and relaxed
//                block-wise: the tile is 8x8 blocks of 8x8 cells, a 64-bit mask says which blocks may still
//                change, a warp takes a dirty block (2 cells per lane), iterates it until it is quiet and
//                marks the neighbouring blocks whose edge it changed.  Work therefore follows the wave fronts
//                instead of sweeping 4096 cells per pass.  A tile whose outer ring changed appends its
//                neighbours to the next round's list (flag + atomic append); grid.sync() separates rounds;
//                the kernel ends when a round's list is empty.  No host round trips.
//   k_nf_verify  one stencil pass checks the equation everywhere.  Relaxed cells satisfy it by
//                construction; a seed can only fail by being too LOW (its lower neighbour is closer than
//                the accumulated epsilons).  Failing seeds are banned and the solve restarts — the result
//                that passes is certified by uniqueness.
//
// CAP (fast path, certified by the same verification): in the solver `sz` holds the plain fill F instead of z
// and candidates above F + capB are ignored.  Every candidate is an upper bound of the solution, the solution
// is >= F >= z so max(., z) never binds on a non-seed cell, and the solution lies within (#non-seed cells) *
// diag of F; a higher candidate (typically from a shore cell millimetres above a lake) can only be provisional
// garbage that the wave from the lake's outlet would overwrite.  Ignoring it is safe for a from-above
// relaxation (the cell just stays at +inf longer) and keeps the work proportional to the lake area.  If the
// verification fails, the solve is repeated without the cap.
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace ms {

constexpr int NF_T = 64;             // tile edge
constexpr int NF_LD = NF_T + 3;      // shared row stride in doubles (odd: spreads the rows of a block over banks)
constexpr int NF_SMEM = (NF_T + 2) * NF_LD * 8 + NF_T * NF_T * 4;
constexpr int NF_BLOCK_ITERS = 64;   // in-block iteration guard (a block that hits it stays dirty)

#ifdef NF_STATS
__device__ unsigned long long g_nf_dbg[4 + 8 * 4096];   // [0] tile iterations [1] block visits [2] block iterations [3] -, then per round (n, ns)
__device__ inline unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#endif

struct NfCtl {
    int count[3];      // entries of tile list k (round r reads list r%3, appends to (r+1)%3, clears (r+2)%3)
    int ticket[3];     // next entry of list k to hand out
    int nonseed;       // cells k_nf_init left at +inf
    int nviol;         // cells failing k_nf_verify
    int rounds;
    int visits;
};

__device__ inline double dmin2(double a, double b) { return a <= b ? a : b; }
__device__ inline double dmin4(double a, double b, double c, double d) { return dmin2(dmin2(a, b), dmin2(c, d)); }

__global__ void __launch_bounds__(256) k_nf_init(const float *__restrict__ z, const float *__restrict__ F,
                                                 double *__restrict__ W, const uint8_t *__restrict__ banned,
                                                 int *tileflag, NfCtl *ctl, int rows, int cols, int tiles_x) {
    int c = blockIdx.x * 64 + (threadIdx.x & 63);
    int r = blockIdx.y * 4 + (threadIdx.x >> 6);
    bool nonseed = false;
    if (r < rows && c < cols) {
        int i = r * cols + c;
        float zc = z[i];
        if (r == 0 || c == 0 || r == rows - 1 || c == cols - 1) {
            W[i] = (double)zc;
        } else {
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
            nonseed = !seed;
        }
    }
    // the 4 rows x 64 columns of this CTA lie in one tile
    int cnt = __syncthreads_count(nonseed);
    if (threadIdx.x == 0 && cnt) {
        atomicAdd(&ctl->nonseed, cnt);
        tileflag[((blockIdx.y * 4) / NF_T) * tiles_x + blockIdx.x] = 16;
    }
}

// flagged tiles -> list 0 (flags stay set: the solver clears a tile's flag when it picks the tile up)
__global__ void __launch_bounds__(256) k_nf_compact(const int *tileflag, int *list, NfCtl *ctl, int ntiles) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    if (tileflag[t]) list[atomicAdd(&ctl->count[0], 1)] = t;
}

// ---- in-tile relaxation -------------------------------------------------------------------------------
// A warp relaxes one 8x8 block until it is quiet.  Lane l owns cells (row l/4, columns 2*(l%4), +1).  block()
// returns whether anything changed; `sides` bit 0/1/2/3 = the block's top / bottom / left / right edge changed,
// bit 4 = the iteration guard was hit (block must stay dirty).
__device__ inline unsigned nf_sides(unsigned bal) {
    unsigned sd = 0;
    if (bal & 0x0000000fu) sd |= 1u;
    if (bal & 0xf0000000u) sd |= 2u;
    if (bal & 0x11111111u) sd |= 4u;
    if (bal & 0x88888888u) sd |= 8u;
    return sd;
}

template <bool CAP>
struct RelaxF64 {
    double *sw;
    const float *sz;
    double sh, dg, capB;
    __device__ inline bool block(int b, unsigned *sides) const {
        const unsigned full = 0xffffffffu;
        int lane = threadIdx.x & 31;
        int lr = (b >> 3) * 8 + (lane >> 2), lc = (b & 7) * 8 + (lane & 3) * 2;
        double *p = sw + (lr + 1) * NF_LD + (lc + 1);
        float2 zz = *reinterpret_cast<const float2 *>(sz + lr * NF_T + lc);
        double z0 = (double)zz.x, z1 = (double)zz.y;
        double w0 = p[0], w1 = p[1];
        bool live = (w0 > z0) || (w1 > z1);
        *sides = 0;
        if (!__any_sync(full, live)) return false;
        unsigned sd = 0;
        bool any = false;
        for (int it = 0;; it++) {
            bool ch = false;
            if (live) {
                double a0 = p[-NF_LD - 1], a1 = p[-NF_LD], a2 = p[-NF_LD + 1], a3 = p[-NF_LD + 2];
                double l = p[-1], r = p[2];
                double c0 = p[NF_LD - 1], c1 = p[NF_LD], c2 = p[NF_LD + 1], c3 = p[NF_LD + 2];
                if (w0 > z0) {
                    double m = dmin2(__dadd_rn(dmin4(a0, a2, c0, c2), dg), __dadd_rn(dmin4(a1, l, w1, c1), sh));
                    if (!CAP) m = m >= z0 ? m : z0;
                    if (m < w0 && (!CAP || m <= z0 + capB)) { w0 = m; p[0] = m; ch = true; }
                }
                if (w1 > z1) {
                    double m = dmin2(__dadd_rn(dmin4(a1, a3, c1, c3), dg), __dadd_rn(dmin4(a2, w0, r, c2), sh));
                    if (!CAP) m = m >= z1 ? m : z1;
                    if (m < w1 && (!CAP || m <= z1 + capB)) { w1 = m; p[1] = m; ch = true; }
                }
            }
            __syncwarp();
            unsigned bal = __ballot_sync(full, ch);
            if (!bal) break;
            any = true;
            sd |= nf_sides(bal);
            if (it == NF_BLOCK_ITERS - 1) { sd |= 16u; break; }
        }
        *sides = sd;
        return any;
    }
};

// Integer form of the same relaxation for a tile whose lake / flat cells all lie in one float64 binade (the
// common case): there W = F + D * ulp with an integer D, `x (+) short` adds exactly sq ulps and `x (+) diag` rounds
// to exactly dq ulps more, so the relaxation is an integer chamfer distance transform.  Cells that
// cannot change (seeds, the raster border, cells outside the raster) are walls; they were taken into account as
// sources once, by k_nf_seedcand.
constexpr int NF_ILD = NF_T + 3;          // shared row stride in ints
constexpr int D_INF = 0x3fffffff;         // lake cell not reached yet
constexpr int D_WALL = 0x40000000;        // not updatable, not a source
constexpr int D_LIMIT = 0x20000000;       // a tile whose distances get this large is left to the float64 form

__device__ inline int imin4(int a, int b, int c, int d) { return min(min(a, b), min(c, d)); }

struct RelaxI32 {
    int *sd;
    int sq, dq;
    int *overflow;     // set when a distance leaves the range the integer form is trusted for
    __device__ inline bool block(int b, unsigned *sides) const {
        const unsigned full = 0xffffffffu;
        int lane = threadIdx.x & 31;
        int lr = (b >> 3) * 8 + (lane >> 2), lc = (b & 7) * 8 + (lane & 3) * 2;
        int *p = sd + (lr + 1) * NF_ILD + (lc + 1);
        int w0 = p[0], w1 = p[1];
        bool live = (w0 <= D_INF) || (w1 <= D_INF);
        *sides = 0;
        if (!__any_sync(full, live)) return false;
        unsigned sds = 0;
        bool any = false;
        for (int it = 0;; it++) {
            bool ch = false;
            if (live) {
                int a0 = p[-NF_ILD - 1], a1 = p[-NF_ILD], a2 = p[-NF_ILD + 1], a3 = p[-NF_ILD + 2];
                int l = p[-1], r = p[2];
                int c0 = p[NF_ILD - 1], c1 = p[NF_ILD], c2 = p[NF_ILD + 1], c3 = p[NF_ILD + 2];
                if (w0 <= D_INF) {
                    int m = min(imin4(a0, a2, c0, c2) + dq, imin4(a1, l, w1, c1) + sq);
                    if (m < w0) { w0 = m; p[0] = m; ch = true; if (m >= D_LIMIT) *overflow = 1; }
                }
                if (w1 <= D_INF) {
                    int m = min(imin4(a1, a3, c1, c3) + dq, imin4(a2, w0, r, c2) + sq);
                    if (m < w1) { w1 = m; p[1] = m; ch = true; if (m >= D_LIMIT) *overflow = 1; }
                }
            }
            __syncwarp();
            unsigned bal = __ballot_sync(full, ch);
            if (!bal) break;
            any = true;
            sds |= nf_sides(bal);
            if (it == NF_BLOCK_ITERS - 1) { sds |= 16u; break; }
        }
        *sides = sds;
        return any;
    }
};

struct NfTileShared {
    unsigned long long dirty[3];
    unsigned long long chgmask;
    int ring;          // bit 0/1/2/3: the tile's top / bottom / left / right ring changed
    int k;             // ticket
    int flags;         // side bits this tile was queued with
    int e;             // common binade exponent of the tile's lake cells (integer form)
    int bad;           // tile does not qualify for the integer form
    int dmax;          // largest finite distance loaded
};

// blocks of the tile that can be affected by what the tile was queued for (bit 4: everything)
__device__ inline unsigned long long nf_region(int flags) {
    if (flags & 16) return ~0ull;
    unsigned long long m = 0;
    if (flags & 1) m |= 0xffull;
    if (flags & 2) m |= 0xffull << 56;
    if (flags & 4) m |= 0x0101010101010101ull;
    if (flags & 8) m |= 0x8080808080808080ull;
    return m;
}

// Runs the dirty-block iteration of one tile to quiescence.  On entry S.dirty[0] holds the initial mask,
// S.dirty[1] = S.dirty[2] = S.chgmask = 0, S.ring = 0, all visible (a __syncthreads() has passed).
template <class R>
__device__ inline int nf_tile_iterate(const R &rx, NfTileShared &S) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int it = 0;
    for (;; it++) {
        const unsigned long long m = S.dirty[it % 3];
        if (m == 0) break;
        if (tid == 0) S.dirty[(it + 2) % 3] = 0;
        unsigned long long mm = m, mark = 0, mine = 0;
        int rings = 0;
        for (int idx = 0; mm; idx++) {
            int b = __ffsll((long long)mm) - 1;
            mm &= mm - 1;
            if ((idx & 7) != warp) continue;
            unsigned sd;
            if (!rx.block(b, &sd)) continue;
            int by = b >> 3, bx = b & 7;
            mine |= 1ull << b;
            if (sd & 16u) mark |= 1ull << b;
            // neighbouring blocks that read a changed edge (a corner cell belongs to both of its edges)
            unsigned long long row3 = ((0x7ull << bx) >> 1) & 0xffull;
            int ylo = by > 0 ? by - 1 : 0, yhi = by < 7 ? by + 1 : 7;
            if (sd & 1u) { if (by > 0) mark |= row3 << ((by - 1) * 8); else rings |= 1; }
            if (sd & 2u) { if (by < 7) mark |= row3 << ((by + 1) * 8); else rings |= 2; }
            if (sd & 4u) {
                if (bx > 0) { for (int y = ylo; y <= yhi; y++) mark |= 1ull << (y * 8 + bx - 1); }
                else rings |= 4;
            }
            if (sd & 8u) {
                if (bx < 7) { for (int y = ylo; y <= yhi; y++) mark |= 1ull << (y * 8 + bx + 1); }
                else rings |= 8;
            }
        }
        if (lane == 0) {
            if (mark) atomicOr(&S.dirty[(it + 1) % 3], mark);
            if (mine) atomicOr(&S.chgmask, mine);
            if (rings) atomicOr(&S.ring, rings);
        }
        __syncthreads();
    }
    return it;
}

// lake cell (w > f) -> integer distance; anything else -> wall.  Tracks the common binade.
__device__ inline int nf_to_int(double w, float f, NfTileShared &S, int &e_seen, int &dmax) {
    double fd = (double)f;
    if (!(w > fd)) return D_WALL;
    long long mag = __double_as_longlong(fd) & 0x7fffffffffffffffll;
    if (fd < 0) mag -= 1;                       // values above a negative F have the smaller magnitude
    int e = (int)(mag >> 52) - 1023;
    if (fd == 0.0 || e < -900) { S.bad = 1; return D_WALL; }
    if (e != e_seen) {
        int old = atomicCAS(&S.e, INT_MIN, e);
        if (old != INT_MIN && old != e) S.bad = 1;
        e_seen = e;
    }
    if (w == INFINITY) return D_INF;
    int ew = (int)((__double_as_longlong(w) & 0x7fffffffffffffffll) >> 52) - 1023;
    if (ew != e) { S.bad = 1; return D_WALL; }
    double inv_ulp = __longlong_as_double((long long)(1023 - (e - 52)) << 52);
    double d = (w - fd) * inv_ulp;              // exact: same binade, both multiples of its ulp
    if (!(d < (double)D_LIMIT)) { S.bad = 1; return D_WALL; }
    int di = (int)d;
    dmax = max(dmax, di);
    return di;
}

template <bool CAP>
__global__ void __launch_bounds__(256) k_nf_solve(const float *__restrict__ zsrc, double *W, int *lists, int *tileflag,
                                                  NfCtl *ctl, int rows, int cols, int tiles_x, int tiles_y, int ntiles,
                                                  double sh, double dg, int max_rounds, int use_int) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *sw = reinterpret_cast<double *>(smem_raw);
    float *sz = reinterpret_cast<float *>(smem_raw + (NF_T + 2) * NF_LD * 8);
    int *sdi = reinterpret_cast<int *>(smem_raw);
    __shared__ NfTileShared S;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const double capB = CAP ? ((double)ctl->nonseed + 16.0) * dg * 1.001 : 0.0;

    for (int round = 0; round < max_rounds; round++) {
        const int cur = round % 3, nxt = (round + 1) % 3, clr = (round + 2) % 3;
        const int n = *(volatile int *)&ctl->count[cur];
        if (n == 0) break;
        if (blockIdx.x == 0 && tid == 0) {
            ctl->count[clr] = 0;
            ctl->ticket[clr] = 0;
            ctl->rounds = round + 1;
            ctl->visits += n;
        }
        const int *list = lists + (size_t)cur * ntiles;
        int *listn = lists + (size_t)nxt * ntiles;
#ifdef NF_STATS
        unsigned long long t_round = gtimer();
#endif
        for (;;) {
            __syncthreads();
            if (tid == 0) S.k = atomicAdd(&ctl->ticket[cur], 1);
            __syncthreads();
            const int k = S.k;
            if (k >= n) break;
            const int t = __ldcg(list + k);
#ifdef NF_STATS
            long long tc0 = clock64(), tc1 = 0, tc2 = 0; int nit = 0;
#endif
            const int ty = t / tiles_x, tx = t - ty * tiles_x;
            const int r0 = ty * NF_T, c0 = tx * NF_T;
            if (tid == 0) {
                // picked up: later changes of a neighbour's ring must queue this tile again
                S.flags = atomicExch(tileflag + t, 0);
                S.dirty[1] = 0;
                S.dirty[2] = 0;
                S.chgmask = 0;
                S.ring = 0;
                S.e = INT_MIN;
                S.bad = 0;
                S.dmax = 0;
            }
            __threadfence();
            __syncthreads();
            bool solved = false;
            if (CAP && use_int) {
                // ---- integer form: load tile + apron, convert on the fly
                int e_seen = INT_MIN, dmax = 0;
                const bool vec = ((cols & 1) == 0) && (c0 + NF_T <= cols);
                for (int lr = warp; lr < NF_T + 2; lr += 8) {
                    int r = r0 + lr - 1;
                    int *row = sdi + lr * NF_ILD;
                    if (r < 0 || r >= rows) {
                        row[1 + 2 * lane] = D_WALL;
                        row[2 + 2 * lane] = D_WALL;
                        if (lane < 2) row[lane ? NF_T + 1 : 0] = D_WALL;
                        continue;
                    }
                    const double *wr = W + (size_t)r * cols;
                    const float *fr = zsrc + (size_t)r * cols;
                    int c = c0 + 2 * lane;
                    double w0v, w1v;
                    float f0v, f1v;
                    if (vec) {
                        double2 wv = __ldcg(reinterpret_cast<const double2 *>(wr + c));
                        float2 fv = __ldg(reinterpret_cast<const float2 *>(fr + c));
                        w0v = wv.x; w1v = wv.y; f0v = fv.x; f1v = fv.y;
                    } else {
                        w0v = c < cols ? __ldcg(wr + c) : 0.0;
                        f0v = c < cols ? __ldg(fr + c) : 0.f;
                        w1v = c + 1 < cols ? __ldcg(wr + c + 1) : 0.0;
                        f1v = c + 1 < cols ? __ldg(fr + c + 1) : 0.f;
                    }
                    row[1 + 2 * lane] = nf_to_int(w0v, f0v, S, e_seen, dmax);
                    row[2 + 2 * lane] = nf_to_int(w1v, f1v, S, e_seen, dmax);
                    if (lane < 2) {
                        int ca = lane ? c0 + NF_T : c0 - 1;
                        int v = D_WALL;
                        if (ca >= 0 && ca < cols) v = nf_to_int(__ldcg(wr + ca), __ldg(fr + ca), S, e_seen, dmax);
                        row[lane ? NF_T + 1 : 0] = v;
                    }
                }
                dmax = __reduce_max_sync(0xffffffffu, dmax);
                if (lane == 0 && dmax) atomicMax(&S.dmax, dmax);
                __syncthreads();
                double ulp = 0, sqd = 0, dqd = 0;
                bool ok = !S.bad;
                if (ok && S.e != INT_MIN) {
                    int e = S.e;
                    ulp = __longlong_as_double((long long)(e - 52 + 1023) << 52);
                    double inv_ulp = __longlong_as_double((long long)(1023 - (e - 52)) << 52);
                    sqd = sh * inv_ulp;
                    double dqx = dg * inv_ulp;
                    dqd = rint(dqx);
                    double fr2 = fabs(dqx - floor(dqx) - 0.5);
                    // short must be a whole number of ulps, diag must not sit on a rounding tie, and the distances
                    // this tile can reach must stay far from the integer range's end
                    ok = sqd >= 1.0 && sqd == rint(sqd) && dqd >= 1.0 && fr2 > 1e-9 && dqd < (double)(1 << 27) &&
                         sqd < (double)(1 << 27) && S.dmax < D_LIMIT;
                }
                if (ok && S.e == INT_MIN) {
                    solved = true;                      // no lake cell in the tile or its apron: nothing to do
                } else if (ok) {
                    if (tid == 0) S.dirty[0] = nf_region(S.flags);
                    __syncthreads();
                    RelaxI32 rx{sdi, (int)sqd, (int)dqd, &S.bad};
#ifdef NF_STATS
                    tc1 = clock64();
#endif
                    int its = nf_tile_iterate(rx, S);
#ifdef NF_STATS
                    nit = its;
#endif
                    (void)its;
                    solved = !S.bad;                    // a distance left the trusted range: redo the tile in float64
                    // write back the blocks that changed: W = F + D * ulp (exact)
                    unsigned long long mm = solved ? S.chgmask : 0ull;
                    for (int idx = 0; mm; idx++) {
                        int b = __ffsll((long long)mm) - 1;
                        mm &= mm - 1;
                        if ((idx & 7) != warp) continue;
                        int lr = (b >> 3) * 8 + (lane >> 2), lc = (b & 7) * 8 + (lane & 3) * 2;
                        int r = r0 + lr, c = c0 + lc;
                        const int *p = sdi + (lr + 1) * NF_ILD + (lc + 1);
                        if (r < rows) {
#pragma unroll
                            for (int q = 0; q < 2; q++) {
                                int d = p[q];
                                if (c + q < cols && d < D_INF) {
                                    size_t i = (size_t)r * cols + c + q;
                                    W[i] = __dadd_rn((double)__ldg(zsrc + i), __dmul_rn((double)d, ulp));
                                }
                            }
                        }
                    }
                }
                if (!solved) __syncthreads();          // everybody is done with the integer tile before it is overwritten
            }
            if (!solved) {
                // ---- float64 form
                if (tid == 0) {
                    S.dirty[0] = nf_region(S.flags);
                    S.dirty[1] = 0;
                    S.dirty[2] = 0;
                    S.chgmask = 0;
                    S.ring = 0;
                }
                for (int q = tid; q < (NF_T + 2) * (NF_T + 2); q += 256) {
                    int lr = q / (NF_T + 2), lc = q - lr * (NF_T + 2);
                    int r = r0 + lr - 1, c = c0 + lc - 1;
                    double v = INFINITY;
                    if (r >= 0 && r < rows && c >= 0 && c < cols) v = __ldcg(W + (size_t)r * cols + c);
                    sw[lr * NF_LD + lc] = v;
                }
                for (int q = tid; q < NF_T * NF_T; q += 256) {
                    int lr = q >> 6, lc = q & 63;
                    int r = r0 + lr, c = c0 + lc;
                    sz[q] = (r < rows && c < cols) ? __ldg(zsrc + (size_t)r * cols + c) : INFINITY;
                }
                __syncthreads();
#ifdef NF_STATS
                tc1 = clock64();
#endif
                RelaxF64<CAP> rx{sw, sz, sh, dg, capB};
                int its = nf_tile_iterate(rx, S);
#ifdef NF_STATS
                nit = its;
#endif
                (void)its;
                // write back the blocks that changed: a warp writes 8 rows of 8 doubles (lane: row l/4, 2 columns)
                unsigned long long mm = S.chgmask;
                for (int idx = 0; mm; idx++) {
                    int b = __ffsll((long long)mm) - 1;
                    mm &= mm - 1;
                    if ((idx & 7) != warp) continue;
                    int lr = (b >> 3) * 8 + (lane >> 2), lc = (b & 7) * 8 + (lane & 3) * 2;
                    int r = r0 + lr, c = c0 + lc;
                    const double *p = sw + (lr + 1) * NF_LD + (lc + 1);
                    if (r < rows) {
                        if (c < cols) W[(size_t)r * cols + c] = p[0];
                        if (c + 1 < cols) W[(size_t)r * cols + c + 1] = p[1];
                    }
                }
            }
#ifdef NF_STATS
            tc2 = clock64();
            if (tid == 0 && round < 4096) {
                atomicMax(&g_nf_dbg[6 + 8 * round], (unsigned long long)(tc1 - tc0));
                atomicMax(&g_nf_dbg[7 + 8 * round], (unsigned long long)(tc2 - tc1));
                atomicMax(&g_nf_dbg[8 + 8 * round], (unsigned long long)nit);
                atomicAdd(&g_nf_dbg[9 + 8 * round], (unsigned long long)(tc2 - tc1));
                atomicAdd(&g_nf_dbg[10 + 8 * round], (unsigned long long)nit);
            }
#endif
            if (S.chgmask == 0) continue;
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                int ring = S.ring;
                bool top = ring & 1, bot = ring & 2, lef = ring & 4, rig = ring & 8;
                for (int dy = -1; dy <= 1; dy++)
                    for (int dx = -1; dx <= 1; dx++) {
                        if (!dy && !dx) continue;
                        // a diagonal neighbour only sees our corner cell: it is covered by either adjoining side
                        bool hit = (dy < 0 && top) || (dy > 0 && bot) || (dx < 0 && lef) || (dx > 0 && rig);
                        if (!hit) continue;
                        int y = ty + dy, x = tx + dx;
                        if (y < 0 || y >= tiles_y || x < 0 || x >= tiles_x) continue;
                        int nb = y * tiles_x + x;
                        // which side of the neighbour looks at us
                        int bits = (dy < 0 ? 2 : 0) | (dy > 0 ? 1 : 0) | (dx < 0 ? 8 : 0) | (dx > 0 ? 4 : 0);
                        if (dy && dx) bits = dy < 0 ? 2 : 1;      // a corner: one block of that side is enough
                        if (atomicOr(tileflag + nb, bits) == 0) listn[atomicAdd(&ctl->count[nxt], 1)] = nb;
                    }
            }
        }
        __threadfence();
        grid.sync();
#ifdef NF_STATS
        if (blockIdx.x == 0 && tid == 0 && round < 4096) {
            g_nf_dbg[4 + 8 * round] = n;
            g_nf_dbg[5 + 8 * round] = gtimer() - t_round;
        }
#endif
    }
}

// CAP only: the one time seeds (and the raster border) act as sources.  A lake / flat cell (still +inf) takes the
// best candidate its fixed neighbours offer, unless that lies above F + capB.  A neighbour is fixed iff W == F
// there (seed: W = z = F; border: W = z = F); lake cells hold +inf or, once written by this kernel, a value > F.
__global__ void __launch_bounds__(256) k_nf_seedcand(const float *__restrict__ F, double *W, const NfCtl *ctl,
                                                     int rows, int cols, double sh, double dg) {
    int c = blockIdx.x * 64 + (threadIdx.x & 63);
    int r = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (r < 1 || c < 1 || r >= rows - 1 || c >= cols - 1) return;
    size_t i = (size_t)r * cols + c;
    if (__ldcg(W + i) != INFINITY) return;
    const double capB = ((double)ctl->nonseed + 16.0) * dg * 1.001;
    double best = INFINITY;
#pragma unroll
    for (int dr = -1; dr <= 1; dr++)
#pragma unroll
        for (int dc = -1; dc <= 1; dc++) {
            if (dr == 0 && dc == 0) continue;
            size_t j = i + (long long)dr * cols + dc;
            double wn = __ldcg(W + j);
            if (wn != (double)__ldg(F + j)) continue;
            double cand = __dadd_rn(wn, (dr != 0 && dc != 0) ? dg : sh);
            best = dmin2(best, cand);
        }
    if (best <= (double)F[i] + capB) W[i] = best;
}

__global__ void __launch_bounds__(256) k_nf_verify(const float *__restrict__ z, const double *__restrict__ W,
                                                   uint8_t *banned, NfCtl *ctl, int rows, int cols, double sh,
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
    if (threadIdx.x == 0 && cnt) atomicAdd(&ctl->nviol, cnt);
}

int g_nf_use_int = 1;      // MS_NF_INT=0 in the environment keeps every tile in the float64 form (debugging)

template <bool CAP>
static int nf_launch_solve(const float *zsrc, double *W, int *lists, int *tileflag, NfCtl *ctl, int rows, int cols,
                           int tiles_x, int tiles_y, int ntiles, double sh, double dg, int max_rounds, int use_int,
                           int64_t units, cudaStream_t s) {
    static int grid_blocks = 0;
    if (!grid_blocks) {
        MS_CUDA(cudaFuncSetAttribute(k_nf_solve<CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, NF_SMEM));
        int dev = 0, sms = 0, per_sm = 0;
        MS_CUDA(cudaGetDevice(&dev));
        MS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        MS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_nf_solve<CAP>, 256, NF_SMEM));
        if (per_sm < 1) { set_error("fill_terrain_no_flats: solver kernel does not fit on an SM"); return MS_ERR_CUDA; }
        grid_blocks = sms * per_sm;
    }
    void *args[] = {(void *)&zsrc, (void *)&W, (void *)&lists, (void *)&tileflag, (void *)&ctl, (void *)&rows,
                    (void *)&cols, (void *)&tiles_x, (void *)&tiles_y, (void *)&ntiles, (void *)&sh, (void *)&dg,
                    (void *)&max_rounds, (void *)&use_int};
    int g = grid_blocks < ntiles ? grid_blocks : ntiles;
    prof_units(units);
    if (g_prof) prof_begin(CAP ? "k_nf_solve<true>" : "k_nf_solve<false>", s);
    cudaError_t e = cudaLaunchCooperativeKernel((const void *)k_nf_solve<CAP>, dim3(g), dim3(256), args, NF_SMEM, s);
    if (g_prof) prof_end(s);
    g_launches++;
    if (e != cudaSuccess) {
        set_error("%s:%d: cooperative launch k_nf_solve -> %s", __FILE__, __LINE__, cudaGetErrorString(e));
        return MS_ERR_CUDA;
    }
    return MS_OK;
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
    static bool env_done = false;
    if (!env_done) {
        const char *e = getenv("MS_NF_INT");
        if (e && e[0] == '0') g_nf_use_int = 0;
        env_done = true;
    }
    int tiles_x = (int)cdiv(cols, NF_T), tiles_y = (int)cdiv(rows, NF_T);
    int ntiles = tiles_x * tiles_y;
    DevBuf<int> tileflag, lists;
    DevBuf<uint8_t> banned;
    DevBuf<NfCtl> ctl;
    MS_TRY(tileflag.alloc((size_t)ntiles, s));
    MS_TRY(lists.alloc((size_t)ntiles * 3, s));
    MS_TRY(ctl.alloc(1, s));
    bool cap = sh > 0 && dg > 0;      // capped fast path first; a verification failure falls back to the generic one
    dim3 g2(cdiv(cols, 64), cdiv(rows, 4));
    NfCtl *h = (NfCtl *)(host_flags().h + 32);
    // every round moves a wave at least one tile; waves can wind, so the bound is generous
    int max_rounds = 64 * (tiles_x + tiles_y) + 1024;
    int64_t rounds = 0, visits = 0, tries = 0;
    for (;;) {
        tries++;
        MS_CUDA(cudaMemsetAsync(tileflag.p, 0, (size_t)ntiles * sizeof(int), s));
        MS_CUDA(cudaMemsetAsync(ctl.p, 0, sizeof(NfCtl), s));
        MS_LAUNCH(k_nf_init, g2, 256, 0, s, dtm, filled, out, banned.p, tileflag.p, ctl.p, (int)rows, (int)cols, tiles_x);
        MS_LAUNCH(k_nf_compact, cdiv(ntiles, 256), 256, 0, s, tileflag.p, lists.p, ctl.p, ntiles);
        if (cap) {
            MS_LAUNCH(k_nf_seedcand, g2, 256, 0, s, filled, out, ctl.p, (int)rows, (int)cols, sh, dg);
            MS_TRY(nf_launch_solve<true>(filled, out, lists.p, tileflag.p, ctl.p, (int)rows, (int)cols, tiles_x, tiles_y,
                                         ntiles, sh, dg, max_rounds, g_nf_use_int, n, s));
        } else {
            MS_TRY(nf_launch_solve<false>(dtm, out, lists.p, tileflag.p, ctl.p, (int)rows, (int)cols, tiles_x, tiles_y,
                                          ntiles, sh, dg, max_rounds, 0, n, s));
        }
        MS_LAUNCH(k_nf_verify, g2, 256, 0, s, dtm, out, banned.p, ctl.p, (int)rows, (int)cols, sh, dg);
        MS_CUDA(cudaMemcpyAsync(h, ctl.p, sizeof(NfCtl), cudaMemcpyDeviceToHost, s));
        MS_TRY(ms::stream_sync(s));
        rounds += h->rounds;
        visits += h->visits;
        if (h->rounds >= max_rounds) {
            set_error("fill_terrain_no_flats: relaxation did not converge in %d rounds", max_rounds);
            return MS_ERR_NOCONV;
        }
        if (h->nviol == 0) break;
        if (cap) {          // the heuristic cap (or a seed) was wrong somewhere: redo without it
            cap = false;
            continue;
        }
        if (!banned.p) {
            // first failure: allocate the ban map and mark the failing seeds
            MS_TRY(banned.alloc((size_t)n, s));
            MS_CUDA(cudaMemsetAsync(banned.p, 0, (size_t)n, s));
            MS_LAUNCH(k_nf_verify, g2, 256, 0, s, dtm, out, banned.p, ctl.p, (int)rows, (int)cols, sh, dg);
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

#ifdef NF_STATS
int ms_nf_debug(unsigned long long *out, int n, int reset) {
    if (out) cudaMemcpyFromSymbol(out, ms::g_nf_dbg, sizeof(unsigned long long) * n);
    if (reset) { static unsigned long long z[4 + 8 * 4096]; cudaMemcpyToSymbol(ms::g_nf_dbg, z, sizeof(z)); }
    return 0;
}
#endif

}  // extern "C"
