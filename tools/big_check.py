"""A raster too large for the CPU oracle: run the device-resident path once, time it, and check the
size-independent certificates with chunked torch arithmetic (test infrastructure, not product code):
  fill        filled >= dem, border filled == dem, and every raised interior cell has no lower filled neighbour
              (a fixed point reached from above is the greatest one)
  no-flats    the library's own verification stencil passed (noflat_reverify == 0 means first try)
  accum       acc(c) == 1 + sum of acc over the cells flowing into c, everywhere; sum over terminal cells == N
  watersheds  ws(c) == label(c) if labelled else ws(downstream(c)) (0 at unlabelled terminals)
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
