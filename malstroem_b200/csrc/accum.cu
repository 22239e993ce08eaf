// accum.cu — K4 flow accumulation, tile-local (shared-memory) tracer + perimeter link forest.
//
// flow.accumulated_flow (flow.py:344-364; speedups/_flow.pyx:225-273): accum(c) = 1 + sum of accum over the
// cells flowing into c = size of c's upstream tree, exact float64 integers.
//
//   pass A  k_acc_tile_a   one CTA per 64x64 tile: D8 codes + 1-cell apron in shared memory, in-tile downstream index
//           per cell, then the size of every cell's in-tile upstream tree by pointer doubling (round 2; round 1 ran
//           the reference's tracer rule from all leaves at once - correct, but a warp sat behind its longest walk).
//           Result: the count each cell collects from inside its own tile (kept as uint16 per cell), and per
//           perimeter cell where its in-tile path ends; cells that leave the tile ("exits") publish their local count.
//   links   k_acc_links / k_acc_node_trace: exits form a forest (exit -> entry cell in the next tile -> the exit
//           that entry's in-tile path ends at).  Same tracer on that forest (~1.5 % of the cells) in global memory
//           gives every exit its full count.
//   pass C  k_acc_tile_c   what enters a tile from the neighbouring exits is added along the in-tile paths below the
//           entry cells (walkers from the <= 252 perimeter cells; the 30-bit inflow is split into two 16-bit halves
//           accumulated in separate 32-bit words, so plain native atomics cannot overflow), and
//           accum = local count + that.
// Codes > 7 and steps off the raster end a path (the reference leaves accumulation over such cells undefined).
#include "common.cuh"

namespace ms {

constexpr int AT = 64;                    // tile edge
constexpr int AH = AT + 2;                // rows of the tile + apron
constexpr int AS = AT + 8;                // shared row stride: columns c0 - 4 .. c0 + 67, so that rows are whole 32-bit words
constexpr int AO = 4;                     // local column lc sits at byte lc + AO of its row
constexpr unsigned short A_OUT = 0xffffu; // "leaves the tile or ends"
constexpr int A_SLOTS = 256;              // perimeter slots per tile (252 used)

__device__ inline int perim_slot(int lr, int lc) {
    if (lr == 0) return lc;
    if (lr == AT - 1) return AT + lc;
    if (lc == 0) return 2 * AT + (lr - 1);
    return 2 * AT + (AT - 2) + (lr - 1);           // lc == AT-1
}
__device__ inline void perim_cell(int p, int *lr, int *lc) {
    if (p < AT) { *lr = 0; *lc = p; }
    else if (p < 2 * AT) { *lr = AT - 1; *lc = p - AT; }
    else if (p < 2 * AT + (AT - 2)) { *lr = p - 2 * AT + 1; *lc = 0; }
    else { *lr = p - (2 * AT + AT - 2) + 1; *lc = AT - 1; }
}

// loads the tile's D8 codes + apron (255 outside the domain) and the in-tile downstream index of every cell
__device__ inline void acc_load_tile(const uint8_t *__restrict__ fd, int rows, int cols, int r0, int c0, int rlo, int rhi,
                                     unsigned char *sdir, unsigned short *sdn) {
    int tid = threadIdx.x;
    if ((cols & 3) == 0 && ((uintptr_t)fd & 3) == 0) {
        // whole 32-bit words: 18 per row instead of 66 byte loads (a word lies entirely inside or outside the raster)
        unsigned *sw = reinterpret_cast<unsigned *>(sdir);
        for (int k = tid; k < AH * (AS / 4); k += 256) {
            int lr = k / (AS / 4), w = k - lr * (AS / 4);
            int r = r0 + lr - 1, c = c0 - AO + 4 * w;
            unsigned v = 0xffffffffu;
            if (r >= rlo && r < rhi && c >= 0 && c < cols)
                v = __ldg(reinterpret_cast<const unsigned *>(fd + (long long)r * cols + c));
            sw[k] = v;
        }
    } else {
        for (int k = tid; k < AH * AH; k += 256) {
            int lr = k / AH, lc = k - lr * AH;
            int r = r0 + lr - 1, c = c0 + lc - 1;
            sdir[lr * AS + lc - 1 + AO] =
                (r >= rlo && r < rhi && c >= 0 && c < cols) ? fd[(long long)r * cols + c] : (unsigned char)255;
        }
    }
    __syncthreads();
    for (int k = tid; k < AT * AT; k += 256) {
        int lr = k >> 6, lc = k & 63;
        int d = sdir[(lr + 1) * AS + (lc + AO)];
        unsigned short dn = A_OUT;
        if (r0 + lr < rows && c0 + lc < cols && d <= 7) {
            int tr = lr + kDR[d], tc = lc + kDC[d];
            if (tr >= 0 && tr < AT && tc >= 0 && tc < AT && r0 + tr < rows && c0 + tc < cols)
                dn = (unsigned short)(tr * AT + tc);
        }
        sdn[k] = dn;
    }
}

// Row bands: `open` bit 0 / 1 = a halo row of flow directions lies above the first / below the last row; a path
// stepping into it leaves the band through an "exit" like any other tile exit.
__global__ void __launch_bounds__(256) k_acc_tile_a(const uint8_t *__restrict__ fd, int rows, int cols, int tiles_x,
                                                    double *__restrict__ nodeX, int *__restrict__ entry_next,
                                                    uint8_t *__restrict__ is_exit, unsigned short *__restrict__ loc16,
                                                    int open) {
    // Round 2: subtree sizes by POINTER DOUBLING instead of walkers from the leaves (whose warps sat behind their
    // longest walk: 18 of 202 ms at 32768^2).  Invariant at the start of round k: S[v] = number of in-tile cells whose
    // path reaches v in fewer than 2^k steps (S = 1: the cell itself), J[v] = the cell exactly 2^k steps downstream
    // (A_OUT when the path has left the tile or ended before).  A round adds S[w] to S[J[w]] for every w - those are
    // the cells reaching J[w] in 2^k .. 2^(k+1) - 1 steps, counted nowhere else - and doubles the jumps; ~log2 of
    // the longest in-tile path rounds, every cell busy in every one.  The additions of a round collect in the high half
    // of the 32-bit word (a count never exceeds 4096), so the low half every other cell reads stays what it was at
    // the start of the round; the new jumps wait in registers until the barrier.  E jumps to the last in-tile cell
    // of the path (for the perimeter cells below).
    __shared__ unsigned int S[AT * AT];
    __shared__ unsigned short J[AT * AT];
    __shared__ unsigned short E[AT * AT];
    __shared__ __align__(16) unsigned char sdir[AH * AS];
    int tile = blockIdx.x;
    int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    int r0 = ty * AT, c0 = tx * AT, tid = threadIdx.x;
    const int rlo = (open & 1) ? -1 : 0, rhi = rows + ((open & 2) ? 1 : 0);
    acc_load_tile(fd, rows, cols, r0, c0, rlo, rhi, sdir, J);
#pragma unroll
    for (int u = 0; u < 16; u++) {
        int k = tid + 256 * u;                      // the cells this thread wrote in acc_load_tile
        unsigned short d = J[k];
        S[k] = 1u;
        E[k] = d == A_OUT ? (unsigned short)k : d;
    }
    __syncthreads();
    for (;;) {
        unsigned short nj[16];
        int any = 0;
        // (batching the loads of 4 or 8 cells ahead of their atomics was tried: 13.6 / 14.0 ms against 12.9 for this
        // plain loop at 32768^2 - the extra registers cost more than the overlapped latency gives)
#pragma unroll
        for (int u = 0; u < 16; u++) {
            int k = tid + 256 * u;
            unsigned short j = J[k], jj = A_OUT;
            if (j != A_OUT) {
                atomicAdd(&S[j], (S[k] & 0xffffu) << 16);
                jj = J[j];
                any = 1;
            }
            nj[u] = jj;
            E[k] = E[E[k]];
        }
        if (!__syncthreads_or(any)) break;
#pragma unroll
        for (int u = 0; u < 16; u++) {
            int k = tid + 256 * u;
            unsigned w = S[k];
            S[k] = (w & 0xffffu) + (w >> 16);
            J[k] = nj[u];
        }
        __syncthreads();
    }
    // (J is A_OUT everywhere now; the in-tile step of a cell is recomputed from its code where needed)
    for (int k = tid; k < AT * AT; k += 256) {
        int lr = k >> 6, lc = k & 63;
        int r = r0 + lr, c = c0 + lc;
        if (r < rows && c < cols) loc16[(size_t)r * cols + c] = (unsigned short)(S[k] & 0xffffu);
    }
    // per perimeter cell, where its in-tile path ends and (for exits) the local count
    if (tid >= 4 * AT - 4) {          // the four unused slots of the tile
        entry_next[(size_t)tile * A_SLOTS + tid] = -1;
        is_exit[(size_t)tile * A_SLOTS + tid] = 0;
    } else {
        int lr, lc;
        perim_cell(tid, &lr, &lc);
        size_t slot = (size_t)tile * A_SLOTS + tid;
        int r = r0 + lr, c = c0 + lc;
        int nxt = -1;
        uint8_t ex = 0;
        if (r < rows && c < cols) {
            const int cur = E[lr * AT + lc];       // the last in-tile cell of the path from here
            // does the end cell step into another tile of the domain?
            int er = cur >> 6, ec = cur & 63;
            int d = sdir[(er + 1) * AS + (ec + AO)];
            if (d <= 7) {
                int gr = r0 + er + kDR[d], gc = c0 + ec + kDC[d];
                if (gr >= rlo && gr < rhi && gc >= 0 && gc < cols) nxt = tile * A_SLOTS + perim_slot(er, ec);
            }
            // is this perimeter cell itself an exit (its own step leaves the tile)?
            int d0 = sdir[(lr + 1) * AS + (lc + AO)];
            if (d0 <= 7 && cur == lr * AT + lc) {
                int gr = r + kDR[d0], gc = c + kDC[d0];
                if (gr >= rlo && gr < rhi && gc >= 0 && gc < cols) {
                    ex = 1;
                    nodeX[slot] = (double)(S[lr * AT + lc] & 0xffffu);
                }
            }
        }
        entry_next[slot] = nxt;
        is_exit[slot] = ex;
    }
}

// (FINAL) halo_top / halo_bot[column] = full count of the neighbouring band's edge-row cell, used where that cell
// flows into this band.
// WIDE = false: one 32-bit word per cell (a count cannot exceed the number of cells of the whole raster; one GPU holds
// at most 2^30) - 16 KB less shared memory, 7 instead of 4 CTAs per SM for a kernel that waits on global loads (ncu:
// long-scoreboard stalls 10.9 per issue, 48 % of the warp slots filled).  WIDE = true (row bands: the raster can have
// 2^32 cells and more): the inflow in two 16-bit halves in separate words.
template <bool WIDE>
__global__ void __launch_bounds__(256) k_acc_tile_c(const uint8_t *__restrict__ fd, int rows, int cols, int tiles_x,
                                                    const double *__restrict__ nodeX,
                                                    const unsigned short *__restrict__ loc16,
                                                    double *__restrict__ accum, int open,
                                                    const double *__restrict__ halo_top,
                                                    const double *__restrict__ halo_bot) {
    __shared__ unsigned int lo[AT * AT], hi[WIDE ? AT * AT : 1];
    __shared__ unsigned short sdn[AT * AT];
    __shared__ __align__(16) unsigned char sdir[AH * AS];
    int tile = blockIdx.x;
    int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    int r0 = ty * AT, c0 = tx * AT, tid = threadIdx.x;
    const int rlo = (open & 1) ? -1 : 0, rhi = rows + ((open & 2) ? 1 : 0);
    for (int k = tid; k < AT * AT; k += 256) { lo[k] = 0; if (WIDE) hi[k] = 0; }
    acc_load_tile(fd, rows, cols, r0, c0, rlo, rhi, sdir, sdn);
    __syncthreads();
    if (tid < 4 * AT - 4) {
        int lr, lc;
        perim_cell(tid, &lr, &lc);
        int r = r0 + lr, c = c0 + lc;
        if (r < rows && c < cols) {
            // what the exits of the neighbouring tiles (or of the neighbouring band) bring into this cell
            const unsigned char *ctr = sdir + (lr + 1) * AS + (lc + AO);
            unsigned long long inflow = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                int nr = lr + kDR[q], nc = lc + kDC[q];
                if (nr >= 0 && nr < AT && nc >= 0 && nc < AT) continue;
                if (ctr[kDR[q] * AS + kDC[q]] != ((q + 4) & 7)) continue;      // apron value 255 never matches
                int gr = r0 + nr, gc = c0 + nc;
                if (gr < 0) inflow += (unsigned long long)halo_top[gc];
                else if (gr >= rows) inflow += (unsigned long long)halo_bot[gc];
                else {
                    int nt = (gr / AT) * tiles_x + (gc / AT);
                    inflow += (unsigned long long)nodeX[(size_t)nt * A_SLOTS + perim_slot(gr % AT, gc % AT)];
                }
            }
            if (inflow) {
                unsigned a = WIDE ? (unsigned)(inflow & 0xffffu) : (unsigned)inflow, b = WIDE ? (unsigned)(inflow >> 16) : 0u;
                int cur = lr * AT + lc;
                for (int guard = 0; guard < AT * AT; guard++) {
                    atomicAdd(&lo[cur], a);
                    if (WIDE && b) atomicAdd(&hi[cur], b);
                    unsigned short d = sdn[cur];
                    if (d == A_OUT) break;
                    cur = d;
                }
            }
        }
    }
    __syncthreads();
    for (int k = tid; k < AT * AT; k += 256) {
        int lr = k >> 6, lc = k & 63;
        int r = r0 + lr, c = c0 + lc;
        if (r < rows && c < cols) {
            size_t i = (size_t)r * cols + c;
            unsigned long long v = (unsigned long long)loc16[i] + lo[k] + (WIDE ? ((unsigned long long)hi[k] << 16) : 0ull);
            accum[i] = (double)v;
        }
    }
}

// exit u -> entry e = down(u) in the neighbouring tile -> the exit e's in-tile path ends at
__global__ void __launch_bounds__(256) k_acc_links(const uint8_t *__restrict__ fd, uint8_t *__restrict__ is_exit,
                                                   const int *__restrict__ entry_next, int *__restrict__ next,
                                                   int *indeg, int rows, int cols, int tiles_x, int nslots) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslots) return;
    int nx = -1;
    if (is_exit[s]) {
        int tile = s / A_SLOTS, p = s - tile * A_SLOTS;
        int lr, lc;
        perim_cell(p, &lr, &lc);
        int r = (tile / tiles_x) * AT + lr, c = (tile % tiles_x) * AT + lc;
        int d = fd[(size_t)r * cols + c];
        int gr = r + kDR[d], gc = c + kDC[d];
        if (gr < 0 || gr >= rows) {
            is_exit[s] = 2;          // leaves the band: a root of the band's link forest (row bands only)
        } else {
            int nt = (gr / AT) * tiles_x + (gc / AT);
            nx = entry_next[(size_t)nt * A_SLOTS + perim_slot(gr % AT, gc % AT)];
            if (nx >= 0) atomicAdd(indeg + nx, 1);
        }
    }
    next[s] = nx;
}

__global__ void __launch_bounds__(256) k_acc_node_trace(const uint8_t *__restrict__ is_exit,
                                                        const int *__restrict__ next, const int *__restrict__ indeg0,
                                                        int *indeg, double *X, int nslots) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nslots || !is_exit[s] || indeg0[s] != 0) return;
    double carried = X[s];
    int cur = s;
    for (;;) {
        int nx = next[cur];
        if (nx < 0) return;
        atomicAdd(X + nx, carried);
        __threadfence();
        if (atomicSub(indeg + nx, 1) != 1) return;
        __threadfence();
        carried = __ldcg(X + nx);
        cur = nx;
    }
}

int accum_dev_impl(const uint8_t *fd, double *acc, int64_t rows, int64_t cols, cudaStream_t s) {
    if (!fd || !acc) { set_error("accumulated_flow: null pointer"); return MS_ERR_ARG; }
    if (rows < 1 || cols < 1 || rows * cols > (1ll << 30)) {
        set_error("accumulated_flow: unsupported shape %lld x %lld", (long long)rows, (long long)cols);
        return MS_ERR_SHAPE;
    }
    int tiles_x = (int)cdiv(cols, AT), tiles_y = (int)cdiv(rows, AT);
    int ntiles = tiles_x * tiles_y;
    int64_t nslots64 = (int64_t)ntiles * A_SLOTS;
    int nslots = (int)nslots64;
    DevBuf<double> X;
    DevBuf<int> entry_next, next, indeg, indeg0;
    DevBuf<uint8_t> is_exit;
    DevBuf<unsigned short> loc16;
    MS_TRY(X.alloc((size_t)nslots, s));
    MS_TRY(entry_next.alloc((size_t)nslots, s));
    MS_TRY(next.alloc((size_t)nslots, s));
    MS_TRY(indeg.alloc((size_t)nslots, s));
    MS_TRY(indeg0.alloc((size_t)nslots, s));
    MS_TRY(is_exit.alloc((size_t)nslots, s));
    MS_TRY(loc16.alloc((size_t)(rows * cols), s));
    MS_CUDA(cudaMemsetAsync(indeg.p, 0, (size_t)nslots * sizeof(int), s));
    MS_CUDA(cudaMemsetAsync(X.p, 0, (size_t)nslots * sizeof(double), s));
    prof_units(rows * cols);
    MS_LAUNCH(k_acc_tile_a, ntiles, 256, 0, s, fd, (int)rows, (int)cols, tiles_x, X.p, entry_next.p, is_exit.p, loc16.p, 0);
    MS_LAUNCH(k_acc_links, cdiv(nslots, 256), 256, 0, s, fd, is_exit.p, entry_next.p, next.p, indeg.p, (int)rows,
              (int)cols, tiles_x, nslots);
    MS_CUDA(cudaMemcpyAsync(indeg0.p, indeg.p, (size_t)nslots * sizeof(int), cudaMemcpyDeviceToDevice, s));
    MS_LAUNCH(k_acc_node_trace, cdiv(nslots, 256), 256, 0, s, is_exit.p, next.p, indeg0.p, indeg.p, X.p, nslots);
    prof_units(rows * cols);
    MS_LAUNCH(k_acc_tile_c<false>, ntiles, 256, 0, s, fd, (int)rows, (int)cols, tiles_x, (const double *)X.p,
              (const unsigned short *)loc16.p, acc, 0, (const double *)nullptr, (const double *)nullptr);
    return MS_OK;
}

// =====================================================================================================
// Row-band accumulation.  Phase 1 (accum_band_local): the tile pass and the band's own link
// forest with nothing flowing in; per band-edge cell it reports (a) where a path leaving the band there enters the
// neighbour and the count it carries, (b) for a cell that receives flow from the neighbour, the band exit its own
// path ends at.  The host side joins these into the forest of band exits of all bands (forest_accumulate).  Phase 2
// (accum_band_finish): the neighbour's totals are injected at the entry cells, the link forest is traced again and
// the tile pass writes the final counts.
// =====================================================================================================
__global__ void __launch_bounds__(256) k_acc_band_out(const uint8_t *__restrict__ fd, const uint8_t *__restrict__ is_exit,
                                                      const double *__restrict__ X, const int *__restrict__ entry_next,
                                                      const int *__restrict__ next, int rows, int cols, int tiles_x,
                                                      int open, int32_t *exit_to, double *exit_val, int32_t *entry_root) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= 2 * cols) return;
    int side = k / cols, c = k - side * cols;
    int to = -1, root = -1;
    double val = 0.0;
    if (open & (1 << side)) {
        int r = side ? rows - 1 : 0;
        int tile = (r / AT) * tiles_x + (c / AT);
        int slot = tile * A_SLOTS + perim_slot(r % AT, c % AT);
        int d = fd[(size_t)r * cols + c];
        if (is_exit[slot] == 2 && d <= 7 && kDR[d] == (side ? 1 : -1)) {
            to = c + kDC[d];
            val = X[slot];
        }
        int s = entry_next[slot];
        if (s >= 0) {
            for (int guard = 0; guard < (1 << 24) && next[s] >= 0; guard++) s = next[s];
            if (is_exit[s] == 2) {
                int t2 = s / A_SLOTS, p = s - t2 * A_SLOTS, lr, lc;
                perim_cell(p, &lr, &lc);
                int rr = (t2 / tiles_x) * AT + lr, cc = (t2 % tiles_x) * AT + lc;
                int d2 = fd[(size_t)rr * cols + cc];
                root = (kDR[d2] > 0 ? cols : 0) + cc;
            }
        }
    }
    exit_to[k] = to;
    exit_val[k] = val;
    entry_root[k] = root;
}

__global__ void __launch_bounds__(256) k_acc_band_inject(const uint8_t *__restrict__ fd, const int *__restrict__ entry_next,
                                                         double *X, int rows, int cols, int tiles_x, int open,
                                                         const double *__restrict__ halo_top,
                                                         const double *__restrict__ halo_bot) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= 2 * cols) return;
    int side = k / cols, c = k - side * cols;
    if (!(open & (1 << side))) return;
    int r = side ? rows - 1 : 0;
    long long hrow = side ? (long long)rows * cols : -(long long)cols;      // the halo row of flow directions
    const double *hT = side ? halo_bot : halo_top;
    double inflow = 0.0;
    for (int dc = -1; dc <= 1; dc++) {
        int cu = c + dc;
        if (cu < 0 || cu >= cols) continue;
        int d = fd[hrow + cu];
        if (d > 7) continue;
        if (kDR[d] == (side ? -1 : 1) && cu + kDC[d] == c) inflow += hT[cu];
    }
    if (inflow == 0.0) return;
    int tile = (r / AT) * tiles_x + (c / AT);
    int s = entry_next[tile * A_SLOTS + perim_slot(r % AT, c % AT)];
    if (s >= 0) atomicAdd(X + s, inflow);
}

__global__ void __launch_bounds__(256) k_fa_indeg(const int32_t *__restrict__ parent, int *indeg, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && parent[i] >= 0) atomicAdd(indeg + parent[i], 1);
}

__global__ void __launch_bounds__(256) k_fa_trace(const int32_t *__restrict__ parent, const int *__restrict__ indeg0,
                                                  int *indeg, double *T, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || indeg0[i] != 0) return;
    double carried = T[i];
    int cur = i;
    for (;;) {
        int nx = parent[cur];
        if (nx < 0) return;
        atomicAdd(T + nx, carried);
        __threadfence();
        if (atomicSub(indeg + nx, 1) != 1) return;
        __threadfence();
        carried = __ldcg(T + nx);
        cur = nx;
    }
}

static int acc_band_bufs(ms_band *B, int *ntiles_out, int *nslots_out) {
    int tiles_x = (int)cdiv(B->cols, AT), tiles_y = (int)cdiv(B->rows, AT);
    int ntiles = tiles_x * tiles_y, nslots = ntiles * A_SLOTS;
    *ntiles_out = ntiles;
    *nslots_out = nslots;
    if (!band_buf(B, BB_ACC_X, (size_t)nslots * 8) || !band_buf(B, BB_ACC_X0, (size_t)nslots * 8) ||
        !band_buf(B, BB_ACC_ENTRY_NEXT, (size_t)nslots * 4) || !band_buf(B, BB_ACC_NEXT, (size_t)nslots * 4) ||
        !band_buf(B, BB_ACC_INDEG, (size_t)nslots * 4) || !band_buf(B, BB_ACC_INDEG0, (size_t)nslots * 4) ||
        !band_buf(B, BB_ACC_ISEXIT, (size_t)nslots) || !band_buf(B, BB_ACC_LOC16, (size_t)(B->rows * B->cols) * 2))
        return MS_ERR_CUDA;
    return MS_OK;
}


}  // namespace ms

extern "C" {

int ms_band_accum_local_dev(ms_band *B, const uint8_t *fd, int32_t *exit_to, double *exit_val, int32_t *entry_root,
                            void *stream) {
    using namespace ms;
    MS_TRY(ensure_init());
    if (!B || !fd || !exit_to || !exit_val || !entry_root) { set_error("band accumulation: null pointer"); return MS_ERR_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    int ntiles, nslots;
    MS_TRY(acc_band_bufs(B, &ntiles, &nslots));
    int rows = (int)B->rows, cols = (int)B->cols, tiles_x = (int)cdiv(cols, AT);
    double *X = (double *)B->buf[BB_ACC_X], *X0 = (double *)B->buf[BB_ACC_X0];
    int *entry_next = (int *)B->buf[BB_ACC_ENTRY_NEXT], *next = (int *)B->buf[BB_ACC_NEXT];
    int *indeg = (int *)B->buf[BB_ACC_INDEG], *indeg0 = (int *)B->buf[BB_ACC_INDEG0];
    uint8_t *is_exit = (uint8_t *)B->buf[BB_ACC_ISEXIT];
    MS_CUDA(cudaMemsetAsync(indeg, 0, (size_t)nslots * sizeof(int), s));
    MS_CUDA(cudaMemsetAsync(X, 0, (size_t)nslots * sizeof(double), s));
    prof_units(B->rows * B->cols);
    MS_LAUNCH(k_acc_tile_a, ntiles, 256, 0, s, fd, rows, cols, tiles_x, X, entry_next, is_exit,
              (unsigned short *)B->buf[BB_ACC_LOC16], B->open);
    MS_LAUNCH(k_acc_links, cdiv(nslots, 256), 256, 0, s, fd, is_exit, entry_next, next, indeg, rows, cols, tiles_x, nslots);
    MS_CUDA(cudaMemcpyAsync(indeg0, indeg, (size_t)nslots * sizeof(int), cudaMemcpyDeviceToDevice, s));
    MS_CUDA(cudaMemcpyAsync(X0, X, (size_t)nslots * sizeof(double), cudaMemcpyDeviceToDevice, s));
    MS_LAUNCH(k_acc_node_trace, cdiv(nslots, 256), 256, 0, s, is_exit, next, indeg0, indeg, X, nslots);
    MS_LAUNCH(k_acc_band_out, cdiv(2 * cols, 256), 256, 0, s, fd, is_exit, X, entry_next, next, rows, cols, tiles_x,
              B->open, exit_to, exit_val, entry_root);
    return MS_OK;
}

/* totals over a forest given by parent indices (-1 = root): T[i] += sum of T over the subtree below i */
int ms_forest_accumulate_dev(int64_t n, const int32_t *parent, double *T, void *stream) {
    using namespace ms;
    MS_TRY(ensure_init());
    if (n < 0 || (n && (!parent || !T))) { set_error("forest_accumulate: bad argument"); return MS_ERR_ARG; }
    if (n == 0) return MS_OK;
    cudaStream_t s = (cudaStream_t)stream;
    DevBuf<int> indeg, indeg0;
    MS_TRY(indeg.alloc((size_t)n, s));
    MS_TRY(indeg0.alloc((size_t)n, s));
    MS_CUDA(cudaMemsetAsync(indeg.p, 0, (size_t)n * sizeof(int), s));
    MS_LAUNCH(k_fa_indeg, cdiv(n, 256), 256, 0, s, parent, indeg.p, (int)n);
    MS_CUDA(cudaMemcpyAsync(indeg0.p, indeg.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToDevice, s));
    MS_LAUNCH(k_fa_trace, cdiv(n, 256), 256, 0, s, parent, indeg0.p, indeg.p, T, (int)n);
    return MS_OK;
}

int ms_band_accum_finish_dev(ms_band *B, const uint8_t *fd, const double *halo_total_top, const double *halo_total_bot,
                             double *accum, void *stream) {
    using namespace ms;
    MS_TRY(ensure_init());
    if (!B || !fd || !accum || !B->buf[BB_ACC_X0]) { set_error("band accumulation: finish before the local phase"); return MS_ERR_ARG; }
    if (((B->open & 1) && !halo_total_top) || ((B->open & 2) && !halo_total_bot)) {
        set_error("band accumulation: halo totals missing for an open band edge");
        return MS_ERR_ARG;
    }
    cudaStream_t s = (cudaStream_t)stream;
    int ntiles, nslots;
    MS_TRY(acc_band_bufs(B, &ntiles, &nslots));
    int rows = (int)B->rows, cols = (int)B->cols, tiles_x = (int)cdiv(cols, AT);
    double *X = (double *)B->buf[BB_ACC_X], *X0 = (double *)B->buf[BB_ACC_X0];
    int *entry_next = (int *)B->buf[BB_ACC_ENTRY_NEXT], *next = (int *)B->buf[BB_ACC_NEXT];
    int *indeg = (int *)B->buf[BB_ACC_INDEG], *indeg0 = (int *)B->buf[BB_ACC_INDEG0];
    uint8_t *is_exit = (uint8_t *)B->buf[BB_ACC_ISEXIT];
    MS_CUDA(cudaMemcpyAsync(X, X0, (size_t)nslots * sizeof(double), cudaMemcpyDeviceToDevice, s));
    MS_CUDA(cudaMemcpyAsync(indeg, indeg0, (size_t)nslots * sizeof(int), cudaMemcpyDeviceToDevice, s));
    if (B->open)
        MS_LAUNCH(k_acc_band_inject, cdiv(2 * cols, 256), 256, 0, s, fd, entry_next, X, rows, cols, tiles_x, B->open,
                  halo_total_top, halo_total_bot);
    MS_LAUNCH(k_acc_node_trace, cdiv(nslots, 256), 256, 0, s, is_exit, next, indeg0, indeg, X, nslots);
    prof_units(B->rows * B->cols);
    MS_LAUNCH(k_acc_tile_c<true>, ntiles, 256, 0, s, fd, rows, cols, tiles_x, (const double *)X,
              (const unsigned short *)B->buf[BB_ACC_LOC16], accum, B->open, halo_total_top, halo_total_bot);
    return MS_OK;
}

}  // extern "C"
