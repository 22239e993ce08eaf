/*
 * ms_oracle.c — CPU restatement of the malstroem raster hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain C, single thread, written from the behaviour of SDFIdk/malstroem (citations are relative to
 * /root/reference).  It is the checker for the CUDA path in malstroem_b200/: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * never links or calls it.
 *
 * Parity pin: tests/test_oracle_golden.py checks every function below against tests/golden/*.npz, which
 * were produced by the reference itself (its compiled Cython modules and its pure-Python path) with
 * tests/golden/make_golden.py, and against the reference's own golden rasters (tests/data/*.tif).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (no -ffast-math: denormals, -0.0 and exact fp64
 * rounding of `x + diag` all matter).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define IDX(r, c) ((size_t)(r) * (size_t)cols + (size_t)(c))

/* ------------------------------------------------------------------------------------------------
 * Plain depression fill.  malstroem/algorithms/fill.py:102-109 (init: +inf interior, border = dtm),
 * :112-171 (UL, LR, UR, LL sweeps until one sweep changes nothing), :34-62 (cell update).
 * Sweeps are CLOSED ranges as in the pure-Python path (fill.py:27-28), which is the semantics of
 * record: the Cython sweep is half-open and can stop one row/column early.
 * Returns the number of sweeps done.
 * ---------------------------------------------------------------------------------------------- */
static int sweep_fill(const float *dtm, float *w, int64_t rows, int64_t cols, int64_t r0, int64_t r1,
                      int64_t c0, int64_t c1)
{
    int64_t rs = r1 > r0 ? 1 : -1, cs = c1 > c0 ? 1 : -1;
    int changed = 0;
    for (int64_t r = r0; r != r1 + rs; r += rs) {
        for (int64_t c = c0; c != c1 + cs; c += cs) {
            float fv = w[IDX(r, c)], z = dtm[IDX(r, c)];
            if (!(fv > z)) continue;
            float m = fv, n;
            n = w[IDX(r - 1, c - 1)]; m = n <= m ? n : m;
            n = w[IDX(r - 1, c)];     m = n <= m ? n : m;
            n = w[IDX(r - 1, c + 1)]; m = n <= m ? n : m;
            n = w[IDX(r, c - 1)];     m = n <= m ? n : m;
            n = w[IDX(r, c + 1)];     m = n <= m ? n : m;
            n = w[IDX(r + 1, c - 1)]; m = n <= m ? n : m;
            n = w[IDX(r + 1, c)];     m = n <= m ? n : m;
            n = w[IDX(r + 1, c + 1)]; m = n <= m ? n : m;
            float nv = m >= z ? m : z;
            if (nv != fv) { w[IDX(r, c)] = nv; changed = 1; }
        }
    }
    return changed;
}

int orc_fill_terrain(const float *dtm, float *w, int64_t rows, int64_t cols)
{
    for (int64_t r = 0; r < rows; r++)
        for (int64_t c = 0; c < cols; c++)
            w[IDX(r, c)] = (r == 0 || c == 0 || r == rows - 1 || c == cols - 1) ? dtm[IDX(r, c)] : INFINITY;
    if (rows < 3 || cols < 3) return 0;
    int64_t mr = rows - 2, mc = cols - 2;
    int n = 0;
    for (;;) {
        n++; if (!sweep_fill(dtm, w, rows, cols, 1, mr, 1, mc)) break;   /* UL */
        n++; if (!sweep_fill(dtm, w, rows, cols, mr, 1, mc, 1)) break;   /* LR */
        n++; if (!sweep_fill(dtm, w, rows, cols, 1, mr, mc, 1)) break;   /* UR */
        n++; if (!sweep_fill(dtm, w, rows, cols, mr, 1, 1, mc)) break;   /* LL */
    }
    return n;
}

/* ------------------------------------------------------------------------------------------------
 * No-flats fill (float64).  fill.py:174-232 driver, :77-99 cell update:
 *   m = min(min(4 diagonal nbrs) + diag, min(4 edge nbrs) + short, self);  new = max(m, dtm)
 * ---------------------------------------------------------------------------------------------- */
static inline double dmin(double a, double b) { return a <= b ? a : b; }

static int sweep_noflat(const float *dtm, double *w, int64_t rows, int64_t cols, int64_t r0, int64_t r1,
                        int64_t c0, int64_t c1, double sh, double dg)
{
    int64_t rs = r1 > r0 ? 1 : -1, cs = c1 > c0 ? 1 : -1;
    int changed = 0;
    for (int64_t r = r0; r != r1 + rs; r += rs) {
        for (int64_t c = c0; c != c1 + cs; c += cs) {
            double fv = w[IDX(r, c)], z = (double)dtm[IDX(r, c)];
            if (!(fv > z)) continue;
            double m = dmin(w[IDX(r - 1, c - 1)],
                            dmin(w[IDX(r - 1, c + 1)], dmin(w[IDX(r + 1, c - 1)], w[IDX(r + 1, c + 1)]))) + dg;
            m = dmin(m, dmin(w[IDX(r - 1, c)],
                             dmin(w[IDX(r, c - 1)], dmin(w[IDX(r, c + 1)], w[IDX(r + 1, c)]))) + sh);
            m = dmin(m, fv);
            double nv = m >= z ? m : z;
            if (nv != fv) { w[IDX(r, c)] = nv; changed = 1; }
        }
    }
    return changed;
}

int orc_fill_terrain_no_flats(const float *dtm, double *w, int64_t rows, int64_t cols, double sh, double dg)
{
    for (int64_t r = 0; r < rows; r++)
        for (int64_t c = 0; c < cols; c++)
            w[IDX(r, c)] = (r == 0 || c == 0 || r == rows - 1 || c == cols - 1) ? (double)dtm[IDX(r, c)] : INFINITY;
    if (rows < 3 || cols < 3) return 0;
    int64_t mr = rows - 2, mc = cols - 2;
    int n = 0;
    for (;;) {
        n++; if (!sweep_noflat(dtm, w, rows, cols, 1, mr, 1, mc, sh, dg)) break;
        n++; if (!sweep_noflat(dtm, w, rows, cols, mr, 1, mc, 1, sh, dg)) break;
        n++; if (!sweep_noflat(dtm, w, rows, cols, 1, mr, mc, 1, sh, dg)) break;
        n++; if (!sweep_noflat(dtm, w, rows, cols, mr, 1, 1, mc, sh, dg)) break;
    }
    return n;
}

/* ------------------------------------------------------------------------------------------------
 * D8 flow direction.  speedups/_flow.pyx:98-176 (the compiled form: diagonals are multiplied by
 * INV_SQRT2 = 1/(2**0.5), strict `>` so the first maximum in the order Up, UpRight, Right, DownRight,
 * Down, DownLeft, Left, UpLeft wins; 8 = no lower neighbour), then flow.py:118-139 for the border.
 * ---------------------------------------------------------------------------------------------- */
static const int DR[8] = {-1, -1, 0, 1, 1, 1, 0, -1};
static const int DC[8] = {0, 1, 1, 1, 0, -1, -1, -1};

void orc_flowdir(const double *t, uint8_t *out, int64_t rows, int64_t cols, int edges_outward)
{
    const double SQRT2 = pow(2.0, 0.5);
    const double INV_SQRT2 = 1.0 / SQRT2;
    for (size_t i = 0; i < (size_t)rows * (size_t)cols; i++) out[i] = 8;
    for (int64_t r = 1; r <= rows - 2; r++) {
        for (int64_t c = 1; c <= cols - 2; c++) {
            double z = t[IDX(r, c)], dzmax = 0.0;
            uint8_t code = 8;
            for (int k = 0; k < 8; k++) {
                double dz = z - t[IDX(r + DR[k], c + DC[k])];
                if (k & 1) dz = dz * INV_SQRT2;
                if (dz > dzmax) { dzmax = dz; code = (uint8_t)k; }
            }
            out[IDX(r, c)] = code;
        }
    }
    if (edges_outward) {
        int64_t mr = rows - 1, mc = cols - 1;
        for (int64_t c = 0; c < cols; c++) out[IDX(0, c)] = 0;
        for (int64_t c = 0; c < cols; c++) out[IDX(mr, c)] = 4;
        for (int64_t r = 0; r < rows; r++) out[IDX(r, 0)] = 6;
        for (int64_t r = 0; r < rows; r++) out[IDX(r, mc)] = 2;
        out[IDX(0, 0)] = 7; out[IDX(0, mc)] = 1; out[IDX(mr, 0)] = 5; out[IDX(mr, mc)] = 3;
    }
}

/* ------------------------------------------------------------------------------------------------
 * Flow accumulation.  speedups/_flow.pyx:225-273: for every cell in raster order run the tracer: sum
 * the accumulation of in-raster neighbours that point here; stop if one of them is still <= 0; else
 * store sum+1, step downstream, repeat until the raster is left.  (mode 0: this re-tracing form,
 * O(N * path); mode 1: the same values by one topological pass over in-degrees, for large rasters.)
 * Codes > 7 do not flow anywhere (_flow.pyx:216-219).  A cell with code > 7 ends the trace (the
 * reference leaves this undefined, _flow.pyx:203; after a no-flats fill no interior cell has it).
 * ---------------------------------------------------------------------------------------------- */
static inline int in_raster(int64_t rows, int64_t cols, int64_t r, int64_t c)
{
    return r >= 0 && r < rows && c >= 0 && c < cols;
}

void orc_accum(const uint8_t *fd, double *acc, int64_t rows, int64_t cols, int mode)
{
    size_t n = (size_t)rows * (size_t)cols;
    memset(acc, 0, n * sizeof(double));
    if (mode == 0) {
        for (int64_t r0 = 0; r0 < rows; r0++)
            for (int64_t c0 = 0; c0 < cols; c0++) {
                int64_t r = r0, c = c0;
                while (in_raster(rows, cols, r, c)) {
                    double s = 0;
                    int unresolved = 0;
                    for (int k = 0; k < 8; k++) {
                        int64_t nr = r + DR[k], nc = c + DC[k];
                        if (!in_raster(rows, cols, nr, nc)) continue;
                        uint8_t d = fd[IDX(nr, nc)];
                        if (d > 7 || ((k + 4) & 7) != d) continue;
                        double ua = acc[IDX(nr, nc)];
                        if (ua <= 0) { unresolved = 1; break; }
                        s += ua;
                    }
                    if (unresolved) break;
                    acc[IDX(r, c)] = s + 1;
                    uint8_t d = fd[IDX(r, c)];
                    if (d > 7) break;
                    r += DR[d]; c += DC[d];
                }
            }
        return;
    }
    /* mode 1: Kahn order */
    uint8_t *indeg = (uint8_t *)calloc(n, 1);
    int64_t *stack = (int64_t *)malloc(n * sizeof(int64_t));
    size_t top = 0;
    for (int64_t r = 0; r < rows; r++)
        for (int64_t c = 0; c < cols; c++) {
            uint8_t d = fd[IDX(r, c)];
            if (d > 7) continue;
            int64_t nr = r + DR[d], nc = c + DC[d];
            if (in_raster(rows, cols, nr, nc)) indeg[IDX(nr, nc)]++;
        }
    for (size_t i = 0; i < n; i++) { acc[i] = 1; if (!indeg[i]) stack[top++] = (int64_t)i; }
    while (top) {
        int64_t i = stack[--top];
        int64_t r = i / cols, c = i % cols;
        uint8_t d = fd[i];
        if (d > 7) continue;
        int64_t nr = r + DR[d], nc = c + DC[d];
        if (!in_raster(rows, cols, nr, nc)) continue;
        size_t j = IDX(nr, nc);
        acc[j] += acc[i];
        if (--indeg[j] == 0) stack[top++] = (int64_t)j;
    }
    /* cells on a cycle (never produced by the pipeline) keep in-degree > 0: the reference leaves them 0 */
    for (size_t i = 0; i < n; i++) if (indeg[i]) acc[i] = 0;
    free(indeg); free(stack);
}

/* ------------------------------------------------------------------------------------------------
 * Local watersheds.  flow.py:398-412 + _raster_utils.py:40-60 (start a walk at every border cell, in
 * the order (0,c),(maxr,c) per column then (r,0),(r,maxc) per inner row) and speedups/_flow.pyx:276-315
 * (explicit stack of (cell, downstream label); an `unassigned` cell takes the label, a labelled cell
 * keeps its own and passes it upstream).  In place on int64 labels (callers widen/narrow).
 * ---------------------------------------------------------------------------------------------- */
typedef struct { int64_t r, c, lbl; } wsitem;

static void ws_from(const uint8_t *fd, int64_t *lab, int64_t rows, int64_t cols, int64_t sr, int64_t sc,
                    int64_t unassigned, wsitem **stk, size_t *cap)
{
    size_t top = 0;
    (*stk)[top++] = (wsitem){sr, sc, unassigned};
    while (top) {
        wsitem it = (*stk)[--top];
        int64_t l = lab[IDX(it.r, it.c)];
        if (l == unassigned) { lab[IDX(it.r, it.c)] = it.lbl; l = it.lbl; }
        for (int k = 0; k < 8; k++) {
            int64_t nr = it.r + DR[k], nc = it.c + DC[k];
            if (!in_raster(rows, cols, nr, nc)) continue;
            uint8_t d = fd[IDX(nr, nc)];
            if (d > 7 || ((k + 4) & 7) != d) continue;
            if (top + 1 >= *cap) { *cap *= 2; *stk = (wsitem *)realloc(*stk, *cap * sizeof(wsitem)); }
            (*stk)[top++] = (wsitem){nr, nc, l};
        }
    }
}

void orc_watersheds(const uint8_t *fd, int64_t *lab, int64_t rows, int64_t cols, int64_t unassigned)
{
    size_t cap = 1 << 16;
    wsitem *stk = (wsitem *)malloc(cap * sizeof(wsitem));
    for (int64_t c = 0; c < cols; c++) {
        ws_from(fd, lab, rows, cols, 0, c, unassigned, &stk, &cap);
        ws_from(fd, lab, rows, cols, rows - 1, c, unassigned, &stk, &cap);
    }
    for (int64_t r = 1; r < rows - 1; r++) {
        ws_from(fd, lab, rows, cols, r, 0, unassigned, &stk, &cap);
        ws_from(fd, lab, rows, cols, r, cols - 1, unassigned, &stk, &cap);
    }
    free(stk);
}

/* ------------------------------------------------------------------------------------------------
 * Connected components.  label.py:19-40 calls scipy.ndimage.label (third-party, unpinned; scipy
 * 1.18.1 here) with the full 3x3 structure.  Published behaviour restated: foreground = value != 0
 * (NaN and denormals are foreground, -0.0 is not), 8-connectivity, int32 labels, components numbered
 * 1..n in the order of their first cell in row-major order.  `fg` is the != 0 mask (uint8).
 * ---------------------------------------------------------------------------------------------- */
static int64_t uf_find(int64_t *p, int64_t x)
{
    while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; }
    return x;
}

int64_t orc_label(const uint8_t *fg, int32_t *out, int64_t rows, int64_t cols)
{
    size_t n = (size_t)rows * (size_t)cols;
    int64_t *p = (int64_t *)malloc(n * sizeof(int64_t));
    for (size_t i = 0; i < n; i++) p[i] = (int64_t)i;
    for (int64_t r = 0; r < rows; r++)
        for (int64_t c = 0; c < cols; c++) {
            if (!fg[IDX(r, c)]) continue;
            const int pr[4] = {0, -1, -1, -1}, pc[4] = {-1, -1, 0, 1};
            for (int k = 0; k < 4; k++) {
                int64_t nr = r + pr[k], nc = c + pc[k];
                if (!in_raster(rows, cols, nr, nc) || !fg[IDX(nr, nc)]) continue;
                int64_t a = uf_find(p, (int64_t)IDX(r, c)), b = uf_find(p, (int64_t)IDX(nr, nc));
                if (a < b) p[b] = a; else if (b < a) p[a] = b;
            }
        }
    int64_t next = 0;
    for (size_t i = 0; i < n; i++) {
        if (!fg[i]) { out[i] = 0; continue; }
        int64_t root = uf_find(p, (int64_t)i);
        if ((size_t)root == i) out[i] = (int32_t)(++next);   /* root is the minimum index: seen first */
        else out[i] = out[root];
    }
    free(p);
    return next;
}

/* ------------------------------------------------------------------------------------------------
 * Per-label tables.  label.py:43-75 / speedups/_label.pyx:68-97 (stats: count, sum accumulated in
 * raster order in float64, strict < / > for min / max), label.py:101-166 / _label.pyx:99-128 (arg-min
 * / arg-max: strict comparison, so the first cell in raster order wins ties; untouched labels keep
 * value=+-inf,row=col=-1), label.py:169-180 (bincount), label.py:78-98 (keep LUT).
 * Tables have nlabels+1 entries.  Return -1 if a label is outside [0, nlabels].
 * ---------------------------------------------------------------------------------------------- */
int orc_label_stats(const double *data, const int64_t *lab, int64_t n, int64_t nlabels, double *mn,
                    double *mx, double *sum, int64_t *cnt)
{
    for (int64_t l = 0; l <= nlabels; l++) { mn[l] = INFINITY; mx[l] = -INFINITY; sum[l] = 0; cnt[l] = 0; }
    for (int64_t i = 0; i < n; i++) {
        int64_t l = lab[i];
        if (l < 0 || l > nlabels) return -1;
        double v = data[i];
        cnt[l]++; sum[l] += v;
        if (v < mn[l]) mn[l] = v;
        if (v > mx[l]) mx[l] = v;
    }
    return 0;
}

int orc_label_extreme_index(const double *data, const int64_t *lab, int64_t rows, int64_t cols,
                            int64_t nlabels, int want_max, double *val, int64_t *row, int64_t *col)
{
    for (int64_t l = 0; l <= nlabels; l++) { val[l] = want_max ? -INFINITY : INFINITY; row[l] = -1; col[l] = -1; }
    for (int64_t r = 0; r < rows; r++)
        for (int64_t c = 0; c < cols; c++) {
            int64_t l = lab[IDX(r, c)];
            if (l < 0 || l > nlabels) return -1;
            double v = data[IDX(r, c)];
            if (want_max ? (v > val[l]) : (v < val[l])) { val[l] = v; row[l] = r; col[l] = c; }
        }
    return 0;
}

int orc_label_count(const int64_t *lab, int64_t n, int64_t nbins, int64_t *cnt)
{
    memset(cnt, 0, (size_t)nbins * sizeof(int64_t));
    for (int64_t i = 0; i < n; i++) {
        if (lab[i] < 0 || lab[i] >= nbins) return -1;
        cnt[lab[i]]++;
    }
    return 0;
}

int orc_keep_labels(const int64_t *lab, int64_t n, const uint8_t *keep, int64_t nkeep, uint8_t *out)
{
    for (int64_t i = 0; i < n; i++) {
        if (lab[i] < 0 || lab[i] >= nkeep) return -1;
        out[i] = keep[lab[i]];
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * net.next_downstream_label / net.pourpoint_network (malstroem/algorithms/net.py:142-192),
 * the walk itself being flow.trace_downstream (flow.py:279-301): yield the cell, read its direction, stop at
 * a no-direction code (flow.py:101-115 returns no delta for codes > 7) or when the next cell is outside the
 * raster.  The start cell is visited first (its label equals the source label, so it never answers).
 * For each of the np start cells: down[k] = first label on the path that differs from the start cell's label
 * and, when has_bg, from the background label; found[k] = 0 when the path ends first (Python None).
 * path_len[k] (optional) = cells yielded up to and including the answering cell — the `geometry` list.
 * Returns -1 when a walk exceeds rows*cols steps (cyclic directions: the reference would never return).
 * ---------------------------------------------------------------------------------------------- */
static const int kDR[8] = {-1, -1, 0, 1, 1, 1, 0, -1};
static const int kDC[8] = {0, 1, 1, 1, 0, -1, -1, -1};

int orc_next_downstream_labels(const uint8_t *fd, const int64_t *lab, int64_t rows, int64_t cols, int64_t np,
                               const int64_t *pp_row, const int64_t *pp_col, int64_t bg, int has_bg,
                               int64_t *down, uint8_t *found, int64_t *path_len, int64_t *path_cells,
                               const int64_t *path_off)
{
    for (int64_t k = 0; k < np; k++) {
        int64_t r = pp_row[k], c = pp_col[k], steps = 0;
        int64_t src = lab[IDX(r, c)];
        found[k] = 0;
        down[k] = 0;
        for (;;) {
            if (r < 0 || r >= rows || c < 0 || c >= cols) break;            /* flow.py:294 cell_in_raster */
            if (path_cells) path_cells[path_off[k] + steps] = r * cols + c;  /* net.py:166-167 */
            steps++;
            int64_t l = lab[IDX(r, c)];
            if (l != src && (!has_bg || l != bg)) {                         /* net.py:168-170 */
                found[k] = 1;
                down[k] = l;
                break;
            }
            int d = fd[IDX(r, c)];
            if (d > 7) break;                                               /* flow.py:296-301 */
            r += kDR[d];
            c += kDC[d];
            if (steps > rows * cols) return -1;
        }
        if (path_len) path_len[k] = steps;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Network.rain_event (malstroem/network.py:75-129) for ne events at once.
 * Nodes are given in insertion order; parent[i] = index of the downstream node, -1 for a root
 * (dstrnodeid None), -2 for an id that is not a node (such a node and everything upstream of it is never
 * reached from a root, network.py:100-111,126-128, so it gets no values: present[i] = 0).
 * Per node and event (network.py:75-98):
 *   rainv  = area * mm * 0.001                       (left to right)
 *   up     = sum(spillv of the upstream nodes, in insertion order)     -- Python sum()
 *   total  = rainv + up;  v = min(total, cap);  spillv = max(0, total - cap)
 *   pctv   = 100.0 * v / cap, None (here NaN) when cap == 0
 * sum_mode 0: plain left-to-right float addition (sum() before CPython 3.12).
 * sum_mode 1: Neumaier compensated addition, which is what sum() does for floats from CPython 3.12 on
 *             (Python/bltinmodule.c, builtin_sum_impl: cs_add per item, `if (c && isfinite(c)) total += c`).
 * Outputs are [ne][n] row-major.  Returns -1 on a cycle among reachable nodes (cannot happen from a root).
 * ---------------------------------------------------------------------------------------------- */
int orc_rain_events(int64_t n, const int64_t *parent, const double *area, const double *cap, int64_t ne,
                    const double *mm, int sum_mode, double *rainv, double *spillv, double *v, double *pctv,
                    uint8_t *present)
{
    int64_t *first = malloc(sizeof(int64_t) * (n + 1)), *next = malloc(sizeof(int64_t) * (n + 1));
    int64_t *last = malloc(sizeof(int64_t) * (n + 1)), *order = malloc(sizeof(int64_t) * (n + 1));
    int64_t *stack = malloc(sizeof(int64_t) * (n + 1));
    for (int64_t i = 0; i < n; i++) { first[i] = -1; last[i] = -1; next[i] = -1; present[i] = 0; }
    for (int64_t i = 0; i < n; i++) {           /* upstream_tree[downstream].append(node), network.py:66-69 */
        int64_t p = parent[i];
        if (p < 0) continue;
        if (first[p] < 0) first[p] = i; else next[last[p]] = i;
        last[p] = i;
    }
    int64_t m = 0;
    for (int64_t root = 0; root < n; root++) {  /* network.py:100-111: any order with children before parents */
        if (parent[root] != -1) continue;
        int64_t sp = 0;
        stack[sp++] = root;
        while (sp) {
            int64_t x = stack[--sp];
            order[m++] = x;
            present[x] = 1;
            for (int64_t ch = first[x]; ch >= 0; ch = next[ch]) stack[sp++] = ch;
        }
    }
    for (int64_t e = 0; e < ne; e++) {
        double *R = rainv + e * n, *S = spillv + e * n, *V = v + e * n, *P = pctv + e * n;
        for (int64_t i = 0; i < n; i++) { R[i] = S[i] = V[i] = P[i] = NAN; }
        for (int64_t k = m - 1; k >= 0; k--) {
            int64_t x = order[k];
            double rv = area[x] * mm[e] * 0.001;
            double up = 0.0, comp = 0.0;
            for (int64_t ch = first[x]; ch >= 0; ch = next[ch]) {
                double y = S[ch];
                if (sum_mode == 0) { up = up + y; continue; }
                double t = up + y;
                if (fabs(up) >= fabs(y)) comp += (up - t) + y; else comp += (y - t) + up;
                up = t;
            }
            if (sum_mode == 1 && comp != 0.0 && isfinite(comp)) up += comp;
            double total = rv + up;
            double filled = cap[x] < total ? cap[x] : total;            /* min(total, cap): first argument unless cap < total */
            double d = total - cap[x];
            R[x] = rv;
            S[x] = d > 0.0 ? d : 0.0;                                   /* max(0, d) */
            V[x] = filled;
            P[x] = cap[x] != 0.0 ? 100.0 * filled / cap[x] : NAN;
        }
    }
    free(first); free(next); free(last); free(order); free(stack);
    return 0;
}
