// accum.cu — K4 flow accumulation, tile-local (shared-memory) tracer + perimeter link forest.
//
// flow.accumulated_flow (flow.py:344-364; speedups/_flow.pyx:225-273): accum(c) = 1 + sum of accum over the
// cells flowing into c = size of c's upstream tree, exact float64 integers.
//
//   pass A  k_acc_tile<false>   one CTA per 64x64 tile: D8 codes + 1-cell apron in shared memory, in-tile
//           downstream index and in-degree per cell, then every in-tile leaf walks downstream with shared-memory
//           atomics (add the carried count to the next cell, decrement its in-degree, continue only as the last
//           missing input — the reference's tracer rule run from all leaves at once).  Result: the count each
//           cell collects from inside its own tile.  For every perimeter cell the tile-local end of its path is
//           chased; cells that leave the tile ("exits") publish their local count.
//   links   k_acc_links / k_acc_node_trace: exits form a forest (exit -> entry cell in the next tile -> the exit
//           that entry's in-tile path ends at).  Same tracer on that forest (~1.5 % of the cells) in global memory
//           gives every exit its full count.
//   pass C  k_acc_tile<true>    the tile pass again with the full counts of the neighbouring exits injected at the
//           entry cells; writes accum.
// Codes > 7 and steps off the raster end a path (the reference leaves accumulation over such cells undefined).
#include "common.cuh"

namespace ms {

constexpr int AT = 64;                    // tile edge
constexpr int AH = AT + 2;                // apron row length
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

struct AccSmem {
    unsigned long long acc[AT * AT];
    int indeg[AT * AT];
    unsigned short dn[AT * AT];
    unsigned char dir[AH * AH];
};

template <bool FINAL>
__global__ void __launch_bounds__(256) k_acc_tile(const uint8_t *__restrict__ fd, int rows, int cols, int tiles_x,
                                                  double *__restrict__ nodeX, int *__restrict__ entry_next,
                                                  uint8_t *__restrict__ is_exit, double *__restrict__ accum) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    AccSmem &S = *reinterpret_cast<AccSmem *>(smem_raw);
    int tile = blockIdx.x;
    int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    int r0 = ty * AT, c0 = tx * AT, tid = threadIdx.x;
    for (int k = tid; k < AH * AH; k += 256) {
        int lr = k / AH, lc = k - lr * AH;
        int r = r0 + lr - 1, c = c0 + lc - 1;
        S.dir[k] = (r >= 0 && r < rows && c >= 0 && c < cols) ? fd[(size_t)r * cols + c] : (unsigned char)255;
    }
    __syncthreads();
    for (int k = tid; k < AT * AT; k += 256) {
        int lr = k >> 6, lc = k & 63;
        int r = r0 + lr, c = c0 + lc;
        const unsigned char *ctr = S.dir + (lr + 1) * AH + (lc + 1);
        int d = *ctr;
        unsigned short dn = A_OUT;
        int n = 0;
        unsigned long long start = 1;
        if (r < rows && c < cols) {
            if (d <= 7) {
                int tr = lr + kDR[d], tc = lc + kDC[d];
                if (tr >= 0 && tr < AT && tc >= 0 && tc < AT && r0 + tr < rows && c0 + tc < cols)
                    dn = (unsigned short)(tr * AT + tc);
            }
#pragma unroll
            for (int q = 0; q < 8; q++) {
                int nr = lr + kDR[q], nc = lc + kDC[q];
                if (ctr[kDR[q] * AH + kDC[q]] != ((q + 4) & 7)) continue;      // apron value 255 never matches
                if (nr >= 0 && nr < AT && nc >= 0 && nc < AT) {
                    n++;
                } else if (FINAL) {
                    // upstream neighbour in another tile: it is an exit there; add its full count
                    int gr = r0 + nr, gc = c0 + nc;
                    int nt = (gr / AT) * tiles_x + (gc / AT);
                    start += (unsigned long long)nodeX[(size_t)nt * A_SLOTS + perim_slot(gr % AT, gc % AT)];
                }
            }
        } else {
            n = 1 << 20;      // outside the raster: never a leaf, never reached
        }
        S.dn[k] = dn;
        S.indeg[k] = n ? n : -1;
        S.acc[k] = start;
    }
    __syncthreads();
    for (int k = tid; k < AT * AT; k += 256) {
        if (S.indeg[k] != -1) continue;
        int cur = k;
        unsigned long long carried = S.acc[k];
        for (;;) {
            unsigned short d = S.dn[cur];
            if (d == A_OUT) break;
            atomicAdd(&S.acc[d], carried);
            __threadfence_block();
            if (atomicSub(&S.indeg[d], 1) != 1) break;
            __threadfence_block();
            carried = *(volatile unsigned long long *)&S.acc[d];
            cur = d;
        }
    }
    __syncthreads();
    if (FINAL) {
        for (int k = tid; k < AT * AT; k += 256) {
            int lr = k >> 6, lc = k & 63;
            int r = r0 + lr, c = c0 + lc;
            if (r < rows && c < cols) accum[(size_t)r * cols + c] = (double)S.acc[k];
        }
        return;
    }
    // pass A epilogue: per perimeter cell, where its in-tile path ends and (for exits) the local count
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
            int cur = lr * AT + lc;
            for (int guard = 0; guard < AT * AT && S.dn[cur] != A_OUT; guard++) cur = S.dn[cur];
            // does the end cell step into another in-raster tile?
            int er = cur >> 6, ec = cur & 63;
            int d = S.dir[(er + 1) * AH + (ec + 1)];
            if (d <= 7 && S.dn[cur] == A_OUT) {
                int gr = r0 + er + kDR[d], gc = c0 + ec + kDC[d];
                if (gr >= 0 && gr < rows && gc >= 0 && gc < cols) nxt = tile * A_SLOTS + perim_slot(er, ec);
            }
            // is this perimeter cell itself an exit?
            int d0 = S.dir[(lr + 1) * AH + (lc + 1)];
            if (d0 <= 7 && S.dn[lr * AT + lc] == A_OUT) {
                int gr = r + kDR[d0], gc = c + kDC[d0];
                if (gr >= 0 && gr < rows && gc >= 0 && gc < cols) {
                    ex = 1;
                    nodeX[slot] = (double)S.acc[lr * AT + lc];
                }
            }
        }
        entry_next[slot] = nxt;
        is_exit[slot] = ex;
    }
}

// exit u -> entry e = down(u) in the neighbouring tile -> the exit e's in-tile path ends at
__global__ void __launch_bounds__(256) k_acc_links(const uint8_t *__restrict__ fd, const uint8_t *__restrict__ is_exit,
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
        int nt = (gr / AT) * tiles_x + (gc / AT);
        nx = entry_next[(size_t)nt * A_SLOTS + perim_slot(gr % AT, gc % AT)];
        if (nx >= 0) atomicAdd(indeg + nx, 1);
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
    static bool attr_done = false;
    if (!attr_done) {
        MS_CUDA(cudaFuncSetAttribute(k_acc_tile<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AccSmem)));
        MS_CUDA(cudaFuncSetAttribute(k_acc_tile<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AccSmem)));
        attr_done = true;
    }
    int tiles_x = (int)cdiv(cols, AT), tiles_y = (int)cdiv(rows, AT);
    int ntiles = tiles_x * tiles_y;
    int64_t nslots64 = (int64_t)ntiles * A_SLOTS;
    int nslots = (int)nslots64;
    DevBuf<double> X;
    DevBuf<int> entry_next, next, indeg, indeg0;
    DevBuf<uint8_t> is_exit;
    MS_TRY(X.alloc((size_t)nslots, s));
    MS_TRY(entry_next.alloc((size_t)nslots, s));
    MS_TRY(next.alloc((size_t)nslots, s));
    MS_TRY(indeg.alloc((size_t)nslots, s));
    MS_TRY(indeg0.alloc((size_t)nslots, s));
    MS_TRY(is_exit.alloc((size_t)nslots, s));
    MS_CUDA(cudaMemsetAsync(indeg.p, 0, (size_t)nslots * sizeof(int), s));
    prof_units(rows * cols);
    MS_LAUNCH(k_acc_tile<false>, ntiles, 256, sizeof(AccSmem), s, fd, (int)rows, (int)cols, tiles_x, X.p, entry_next.p,
              is_exit.p, (double *)nullptr);
    MS_LAUNCH(k_acc_links, cdiv(nslots, 256), 256, 0, s, fd, is_exit.p, entry_next.p, next.p, indeg.p, (int)rows,
              (int)cols, tiles_x, nslots);
    MS_CUDA(cudaMemcpyAsync(indeg0.p, indeg.p, (size_t)nslots * sizeof(int), cudaMemcpyDeviceToDevice, s));
    MS_LAUNCH(k_acc_node_trace, cdiv(nslots, 256), 256, 0, s, is_exit.p, next.p, indeg0.p, indeg.p, X.p, nslots);
    prof_units(rows * cols);
    MS_LAUNCH(k_acc_tile<true>, ntiles, 256, sizeof(AccSmem), s, fd, (int)rows, (int)cols, tiles_x, X.p, entry_next.p,
              is_exit.p, acc);
    return MS_OK;
}

}  // namespace ms
