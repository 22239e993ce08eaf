// tile-order simulator for the no-flats chamfer relaxation (CPU model of k_nf_solve_ir's tile FIFO)
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#define T 64
#define INF 0x3fffffff
#define WALL 0x7fffffff
typedef struct { int key, tile; } Ent;
static int rows, cols, tx, ty;
static int32_t *D;
static int SQ = 1024, DQ = 1448;
static inline int32_t get(int r, int c) { return (r < 0 || r >= rows || c < 0 || c >= cols) ? WALL : D[(size_t)r * cols + c]; }
// relax tile to convergence; returns 1 if anything changed; edge-change flags out
static int relax(int t, int *nsweeps) {
    int r0 = (t / tx) * T, c0 = (t % tx) * T, any = 0;
    for (;;) {
        int ch = 0;
        for (int pass = 0; pass < 2; pass++) {
            for (int a = 0; a < T; a++) for (int b = 0; b < T; b++) {
                int lr = pass ? T - 1 - a : a, lc = pass ? T - 1 - b : b;
                int r = r0 + lr, c = c0 + lc;
                if (r >= rows || c >= cols) continue;
                int32_t d = D[(size_t)r * cols + c];
                if (d == WALL) continue;
                int32_t best = d;
                for (int dr = -1; dr <= 1; dr++) for (int dc = -1; dc <= 1; dc++) {
                    if (!dr && !dc) continue;
                    int32_t n = get(r + dr, c + dc);
                    if (n >= INF) continue;
                    int32_t cand = n + ((dr && dc) ? DQ : SQ);
                    if (cand < best) best = cand;
                }
                if (best < d) { D[(size_t)r * cols + c] = best; ch = 1; }
            }
        }
        (*nsweeps)++;
        if (!ch) break;
        any = 1;
    }
    return any;
}
// can neighbour tile nt gain from tile t's current edge? returns min offered distance or -1
static int offer(int t, int dy, int dx) {
    int r0 = (t / tx) * T, c0 = (t % tx) * T;
    int best = -1;
    int rl = dy < 0 ? 0 : (dy > 0 ? T - 1 : 0), rh = dy < 0 ? 0 : (dy > 0 ? T - 1 : T - 1);
    int cl = dx < 0 ? 0 : (dx > 0 ? T - 1 : 0), ch = dx < 0 ? 0 : (dx > 0 ? T - 1 : T - 1);
    for (int lr = rl; lr <= rh; lr++) for (int lc = cl; lc <= ch; lc++) {
        int r = r0 + lr, c = c0 + lc;
        int32_t d = get(r, c);
        if (d >= INF) continue;
        for (int dr = -1; dr <= 1; dr++) for (int dc = -1; dc <= 1; dc++) {
            if (!dr && !dc) continue;
            int rr = r + dr, cc = c + dc;
            // target must lie in the neighbour tile (dy, dx)
            int tyy = (rr < 0 ? -1 : rr / T) - (r0 / T), txx = (cc < 0 ? -1 : cc / T) - (c0 / T);
            if (tyy != dy || txx != dx) continue;
            int32_t n = get(rr, cc);
            if (n == WALL) continue;
            int32_t cand = d + ((dr && dc) ? DQ : SQ);
            if (cand < n && (best < 0 || cand < best)) best = cand;
        }
    }
    return best;
}
// mode 0: FIFO; 1: priority by smallest offered distance (binary heap); 2: bucketed FIFO (key >> shift)
// generation-parallel FIFO: P tiles are "in flight": emulate by processing in batches of P from the queue front, pushes appended after the batch
long simulate(int32_t *Dio, int r, int c, int mode, int shift, int P, long *out_sweeps, long *out_batches) {
    rows = r; cols = c; D = Dio; tx = (cols + T - 1) / T; ty = (rows + T - 1) / T;
    int nt = tx * ty;
    char *queued = calloc(nt, 1);
    int *key = malloc(nt * sizeof(int));
    int cap = nt * 64;
    Ent *q = malloc((size_t)cap * sizeof(Ent));
    long qh = 0, qt = 0, visits = 0, sweeps = 0, batches = 0;
    // initial: every tile with a non-wall cell
    for (int t = 0; t < nt; t++) {
        int r0 = (t / tx) * T, c0 = (t % tx) * T, has = 0;
        for (int lr = 0; lr < T && !has; lr++) for (int lc = 0; lc < T; lc++) if (get(r0 + lr, c0 + lc) != WALL) { has = 1; break; }
        if (has) { q[qt].key = 0; q[qt].tile = t; qt++; queued[t] = 1; key[t] = 0; }
    }
    Ent *batch = malloc((size_t)P * sizeof(Ent));
    while (qh < qt) {
        int nb = 0;
        if (mode == 0) {
            while (qh < qt && nb < P) batch[nb++] = q[qh++];
        } else {
            // take the P smallest keys (mode 1) / entries of the lowest bucket in FIFO order (mode 2); O(n) scan: fine for a model
            for (; nb < P && qh < qt;) {
                long bi = -1; int bk = 0x7fffffff;
                for (long i = qh; i < qt; i++) {
                    if (q[i].tile < 0) continue;
                    int k = mode == 1 ? key[q[i].tile] : (key[q[i].tile] >> shift);
                    if (k < bk) { bk = k; bi = i; }
                }
                if (bi < 0) { qh = qt; break; }
                batch[nb++] = q[bi];
                q[bi].tile = -1;
                while (qh < qt && q[qh].tile < 0) qh++;
            }
        }
        if (!nb) break;
        batches++;
        for (int i = 0; i < nb; i++) queued[batch[i].tile] = 0;
        for (int i = 0; i < nb; i++) {
            int t = batch[i].tile, ns = 0;
            visits++;
            relax(t, &ns);
            sweeps += ns;
            int y = t / tx, x = t % tx;
            for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) {
                if (!dy && !dx) continue;
                int yy = y + dy, xx = x + dx;
                if (yy < 0 || yy >= ty || xx < 0 || xx >= tx) continue;
                int o = offer(t, dy, dx);
                if (o < 0) continue;
                int n = yy * tx + xx;
                if (!queued[n]) {
                    if (qt >= cap) { cap *= 2; q = realloc(q, (size_t)cap * sizeof(Ent)); }
                    q[qt].key = o; q[qt].tile = n; qt++; queued[n] = 1; key[n] = o;
                } else if (o < key[n]) key[n] = o;
            }
        }
    }
    *out_sweeps = sweeps; *out_batches = batches;
    free(queued); free(key); free(q); free(batch);
    return visits;
}
