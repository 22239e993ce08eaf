"""A raster too large for the CPU oracle: run the device-resident path once, time it, and check the
size-independent certificates with chunked torch arithmetic (test / measurement infrastructure, not product code;
none of it calls the library):
  fill        filled >= dem, border filled == dem, and every raised interior cell has no lower filled neighbour
              (a fixed point reached from above is the greatest one)                                   [A.1]
  no-flats    W == max(z, min(min4diag W + diag, min4edge W + short)) at every interior cell, W == z on the border,
              recomputed here in torch float64 (unique fixed point => certifies the surface)             [A.2]
  D8          the flow directions recomputed from the no-flats surface in torch (Up..UpLeft, strict >, diagonals
              times 1/2**0.5, edges flowing outward)
  accum       acc(c) == 1 + sum of acc over the cells flowing into c, everywhere; sum over terminal cells == N
  watersheds  ws(c) == label(c) if labelled else ws(downstream(c)) (0 at unlabelled terminals)
  CC          label != 0 <=> depth != 0; 8-adjacent wet cells share a label; every label 1..n is used and the labels
              increase with the smallest row-major index of their cells (scipy's numbering)
  tables      label_stats / label_count / label_min_index / label_max_index recomputed with torch scatter
              reductions (another schedule): min, max, counts, arg cells exact, sums to 1e-6 relative
usage: python tools/big_check.py [S] [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from malstroem_b200.pipeline import RasterPipeline, synth_fractal

DR = (-1, -1, 0, 1, 1, 1, 0, -1)
DC = (0, 1, 1, 1, 0, -1, -1, -1)


def certify(p, CH=2048, skip_top=0, skip_bottom=0):
    """Certificates on a finished RasterPipeline (or on one band of a BandPipeline: then the first / last own row,
    whose neighbours live in another band, is skipped with skip_top / skip_bottom = 1 and the terminal sum has no
    meaning).  Returns (fill violations, accumulation violations, watershed violations, terminal sum)."""
    R, C = p.rows, p.cols
    dem, filled, acc, fd, lab, ws = p.dem, p.out["filled"], p.out["accum"], p.out["flowdir"], p.out["labels"], p.out["wsheds"]
    dev = dem.device
    bad_fill = bad_acc = bad_ws = 0
    root_sum = 0.0
    cc = torch.arange(C, device=dev).view(1, -1)
    for r0 in range(0, R, CH):
        r1 = min(R, r0 + CH)
        a0, a1 = max(r0 - 1, 0), min(r1 + 1, R)
        F, Z, A, D = filled[a0:a1], dem[a0:a1], acc[a0:a1], fd[a0:a1].long()
        Wl, Lb = ws[a0:a1], lab[a0:a1]
        o0, o1 = r0 - a0, r0 - a0 + (r1 - r0)                   # own rows inside the padded chunk
        Fo, Zo, Ao, Do, Wo, Lo = F[o0:o1], Z[o0:o1], A[o0:o1], D[o0:o1], Wl[o0:o1], Lb[o0:o1]
        rr = torch.arange(r0, r1, device=dev).view(-1, 1)
        keep = ((rr >= skip_top) & (rr < R - skip_bottom)).expand(r1 - r0, C)
        border = (rr == 0) | (rr == R - 1) | (cc == 0) | (cc == C - 1)
        if skip_top:
            border = border & ~((rr == 0) & (cc > 0) & (cc < C - 1))
        if skip_bottom:
            border = border & ~((rr == R - 1) & (cc > 0) & (cc < C - 1))
        bad_fill += int(((Fo < Zo) & keep).sum())
        bad_fill += int(((Fo != Zo) & border & keep).sum())
        upstream = torch.ones_like(Ao)
        wsdown = torch.zeros_like(Wo)
        moves = torch.zeros_like(keep)
        for q in range(8):
            # the neighbour in direction q of every own cell, where it exists inside the padded chunk
            rs, cs = DR[q], DC[q]
            src_r0, src_r1 = max(o0 + rs, 0), min(o1 + rs, F.shape[0])
            if src_r1 <= src_r0:
                continue
            dst_r0 = src_r0 - (o0 + rs)
            dst_r1 = dst_r0 + (src_r1 - src_r0)
            own = (slice(dst_r0, dst_r1), slice(max(-cs, 0), C + min(-cs, 0)))
            nb = (slice(src_r0, src_r1), slice(max(cs, 0), C + min(cs, 0)))
            raised = (Fo[own] > Zo[own]) & ~border[own] & keep[own]
            bad_fill += int((raised & (F[nb] < Fo[own])).sum())
            into = D[nb] == ((q + 4) & 7)                              # that neighbour flows into me
            upstream[own] += torch.where(into, A[nb], torch.zeros_like(upstream[own]))
            mine = Do[own] == q                                        # I flow into that neighbour
            wsdown[own] = torch.where(mine, Wl[nb], wsdown[own])
            moves[own] |= mine
        bad_acc += int(((upstream != Ao) & keep).sum())
        root_sum += float(Ao[~moves].sum())
        want = torch.where(Lo != 0, Lo, torch.where(moves, wsdown, torch.zeros_like(wsdown)))
        bad_ws += int(((want != Wo) & keep).sum())
    return bad_fill, bad_acc, bad_ws, root_sum


INV_SQRT2 = 1.0 / 2 ** 0.5          # _flow.pyx:93-94


def _shift(a, o0, o1, rs, cs, fill):
    """The neighbour (row + rs, col + cs) of the own rows [o0, o1) of the padded chunk `a`; `fill` where there is none."""
    C = a.shape[1]
    out = torch.full((o1 - o0, C), fill, dtype=a.dtype, device=a.device)
    src_r0, src_r1 = max(o0 + rs, 0), min(o1 + rs, a.shape[0])
    if src_r1 <= src_r0:
        return out
    dst_r0 = src_r0 - (o0 + rs)
    out[dst_r0:dst_r0 + (src_r1 - src_r0), max(-cs, 0):C + min(-cs, 0)] = a[src_r0:src_r1, max(cs, 0):C + min(cs, 0)]
    return out


def certify_more(p, short, diag, CH=1024, skip_top=0, skip_bottom=0, row_offset=0, total_rows=None, reduce=None,
                 tables=None, nlabels=None):
    """The certificates big_check.certify does not cover: no-flats stencil, D8 recompute, CC, per-label tables.
    p: RasterPipeline, or one band of a BandPipeline (then row_offset / total_rows place it in the whole raster, the
    band's first / last own row is skipped where its neighbours live in another band, and reduce(t, op) combines the
    per-label partials of all bands: op in 'min' / 'max' / 'sum').  Returns a dict of violation counts."""
    R, C = p.rows, p.cols
    TR = total_rows if total_rows is not None else R
    dev = p.dem.device
    dem, W, fd, acc = p.dem, p.out["fnf"], p.out["flowdir"], p.out["accum"]
    dep, lab, ws = p.out["depths"], p.out["labels"], p.out["wsheds"]
    n = int(nlabels if nlabels is not None else p.nlabels)
    tb = tables if tables is not None else {k: p.table(k) for k in p.tables}
    I64MAX = torch.iinfo(torch.int64).max
    first = torch.full((n + 1,), I64MAX, dtype=torch.int64, device=dev)
    t_min = torch.full((n + 1,), float("inf"), dtype=torch.float64, device=dev)
    t_max = torch.full((n + 1,), float("-inf"), dtype=torch.float64, device=dev)
    t_sum = torch.zeros((n + 1,), dtype=torch.float64, device=dev)
    t_cnt = torch.zeros((n + 1,), dtype=torch.int64, device=dev)
    w_cnt = torch.zeros((n + 1,), dtype=torch.int64, device=dev)
    pmin = torch.full((n + 1,), float("inf"), dtype=torch.float64, device=dev)
    pmax = torch.full((n + 1,), float("-inf"), dtype=torch.float64, device=dev)
    bad = {"noflats": 0, "d8": 0, "cc_adjacent": 0, "cc_foreground": 0, "label_range": 0}
    cc = torch.arange(C, device=dev).view(1, -1)
    chunks = [(r0, min(R, r0 + CH)) for r0 in range(0, R, CH)]
    for r0, r1 in chunks:
        a0, a1 = max(r0 - 1, 0), min(r1 + 1, R)
        o0, o1 = r0 - a0, r0 - a0 + (r1 - r0)
        Wp, Zo, Do = W[a0:a1], dem[r0:r1].double(), fd[r0:r1]
        Wo = Wp[o0:o1]
        rr = torch.arange(r0, r1, device=dev).view(-1, 1)
        keep = ((rr >= skip_top) & (rr < R - skip_bottom)).expand(r1 - r0, C)
        gr = rr + row_offset
        top, bot = gr == 0, gr == TR - 1
        border = top | bot | (cc == 0) | (cc == C - 1)
        inf = float("inf")
        nb = [_shift(Wp, o0, o1, DR[q], DC[q], inf) for q in range(8)]
        d4 = torch.minimum(torch.minimum(nb[1], nb[3]), torch.minimum(nb[5], nb[7]))
        e4 = torch.minimum(torch.minimum(nb[0], nb[2]), torch.minimum(nb[4], nb[6]))
        G = torch.maximum(Zo, torch.minimum(d4 + diag, e4 + short))
        bad["noflats"] += int((torch.where(border, Wo != Zo, Wo != G) & keep).sum())
        code = torch.full(Wo.shape, 8, dtype=torch.uint8, device=dev)
        dzmax = torch.zeros_like(Wo)
        for q in range(8):
            dz = Wo - nb[q]
            if q & 1:
                dz = dz * INV_SQRT2
            up = dz > dzmax
            dzmax = torch.where(up, dz, dzmax)
            code = torch.where(up, torch.full_like(code, q), code)
        del nb, d4, e4, G, dzmax
        left, right = (cc == 0).expand_as(code), (cc == C - 1).expand_as(code)
        for cond, val in ((top.expand_as(code), 0), (bot.expand_as(code), 4), (left, 6), (right, 2), (top & left, 7),
                          (top & right, 1), (bot & left, 5), (bot & right, 3)):
            code = torch.where(cond, torch.full_like(code, val), code)
        bad["d8"] += int(((code != Do) & keep).sum())
        del code
        # connected components
        Lp, Po = lab[a0:a1], dep[r0:r1]
        Lo = Lp[o0:o1]
        fg = Po != 0
        bad["cc_foreground"] += int(((fg != (Lo != 0)) & keep).sum())
        bad["label_range"] += int(((Lo < 0) | (Lo > n)).sum())
        for q in (2, 3, 4, 5):                  # each adjacent pair once: right, down-right, down, down-left
            Ln = _shift(Lp, o0, o1, DR[q], DC[q], 0)
            # (a neighbour outside the padded chunk reads as 0: pairs across a band edge are not checked here)
            bad["cc_adjacent"] += int(((Lo != 0) & (Ln != 0) & (Lo != Ln)).sum())
        idx = ((gr * C) + cc).reshape(-1)
        lf = Lo.reshape(-1).long().clamp(0, n)
        first.scatter_reduce_(0, lf, idx, "amin")
        pv = Po.reshape(-1).double()
        t_min.scatter_reduce_(0, lf, pv, "amin")
        t_max.scatter_reduce_(0, lf, pv, "amax")
        t_sum.scatter_add_(0, lf, pv)
        t_cnt += torch.bincount(lf, minlength=n + 1)
        w_cnt += torch.bincount(ws[r0:r1].reshape(-1).long().clamp(0, n), minlength=n + 1)
        pmin.scatter_reduce_(0, lf, Wo.reshape(-1), "amin")
        pmax.scatter_reduce_(0, lf, acc[r0:r1].reshape(-1), "amax")
    if reduce is not None:
        for t, op in ((first, "min"), (t_min, "min"), (t_max, "max"), (t_sum, "sum"), (t_cnt, "sum"), (w_cnt, "sum"),
                      (pmin, "min"), (pmax, "max")):
            reduce(t, op)
    # second pass: the first cell in raster order that holds the extreme
    imin = torch.full((n + 1,), I64MAX, dtype=torch.int64, device=dev)
    imax = torch.full((n + 1,), I64MAX, dtype=torch.int64, device=dev)
    for r0, r1 in chunks:
        rr = torch.arange(r0, r1, device=dev).view(-1, 1) + row_offset
        idx = ((rr * C) + cc).reshape(-1)
        lf = lab[r0:r1].reshape(-1).long().clamp(0, n)
        wv, av = W[r0:r1].reshape(-1), acc[r0:r1].reshape(-1)
        imin.scatter_reduce_(0, lf, torch.where(wv == pmin[lf], idx, torch.full_like(idx, I64MAX)), "amin")
        imax.scatter_reduce_(0, lf, torch.where(av == pmax[lf], idx, torch.full_like(idx, I64MAX)), "amin")
    if reduce is not None:
        reduce(imin, "min")
        reduce(imax, "min")
    bad["cc_unused_labels"] = int((first[1:] == I64MAX).sum())
    bad["cc_order"] = int((first[1:][1:] <= first[1:][:-1]).sum()) if n > 1 else 0
    m = n + 1
    tv = {}
    tv["st_min"] = int((tb["st_min"][:m] != t_min).sum())
    tv["st_max"] = int((tb["st_max"][:m] != t_max).sum())
    tv["st_count"] = int((tb["st_count"][:m] != t_cnt).sum())
    tv["ws_count"] = int((tb["ws_count"][:m] != w_cnt).sum())
    tol = 1e-6 * torch.maximum(t_sum.abs(), torch.full_like(t_sum, 1e-300))
    tv["st_sum"] = int(((tb["st_sum"][:m] - t_sum).abs() > tol).sum())
    for key, val, ii in (("ppmin", pmin, imin), ("ppmax", pmax, imax)):
        none = ii == I64MAX
        row = torch.where(none, torch.full_like(ii, -1), ii // C)
        col = torch.where(none, torch.full_like(ii, -1), ii % C)
        tv[key] = int(((tb[key + "_value"][:m] != val) | (tb[key + "_row"][:m] != row) | (tb[key + "_col"][:m] != col)).sum())
    bad["tables"] = tv
    return bad


def raster_checksums(p, row_offset=0):
    """Position-weighted 64-bit sums (wrap-around) of the bits of every output raster — additive over row bands, so the
    sum over all bands of a banded run equals the single-GPU value iff (up to hash collisions) every cell is equal."""
    out = {}
    C = p.cols
    for name, t in p.out.items():
        acc_ = torch.zeros((), dtype=torch.int64, device=t.device)
        for r0 in range(0, p.rows, 2048):
            r1 = min(p.rows, r0 + 2048)
            x = t[r0:r1]
            if x.dtype == torch.float64:
                bits = (x + 0.0).view(torch.int64)
            elif x.dtype == torch.float32:
                bits = (x + 0.0).view(torch.int32).long()
            else:
                bits = x.long()
            rr = torch.arange(r0, r1, device=t.device, dtype=torch.int64).view(-1, 1) + row_offset
            idx = rr * C + torch.arange(C, device=t.device, dtype=torch.int64).view(1, -1)
            acc_ += ((bits + 1) * (idx * -7046029254386353131 + 1442695040888963407)).sum()
        out[name] = acc_
    return out


if __name__ == "__main__":
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    p = RasterPipeline(S, S)
    synth_fractal(S, S, seed=1, out=p.dem)
    torch.cuda.synchronize()
    for k in range(reps):
        t0 = time.perf_counter(); p.run(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print("run %d: %.1f ms  %.2f Gcell/s  nlabels %d  stats %s" % (k, dt * 1e3, S * S / dt / 1e9, p.nlabels, p.stats), flush=True)
    bad_fill, bad_acc, bad_ws, root_sum = certify(p)
    print("certificates: fill violations %d, accumulation violations %d (terminal sum %.0f vs N %d), watershed violations %d"
          % (bad_fill, bad_acc, root_sum, S * S, bad_ws))
    print("RESULT", "OK" if (bad_fill == 0 and bad_acc == 0 and bad_ws == 0 and root_sum == S * S) else "FAIL")
