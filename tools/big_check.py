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


def certify(p, CH=2048):
  """Certificates on a finished RasterPipeline `p` (square or not): returns (fill, accum, watershed violations, terminal sum)."""
  S_rows, S = p.rows, p.cols
  dem, filled, acc, fd, lab, ws = p.dem, p.out["filled"], p.out["accum"], p.out["flowdir"], p.out["labels"], p.out["wsheds"]
  bad_fill = bad_acc = bad_ws = 0
  root_sum = 0.0
  for r0 in range(0, S_rows, CH):
      r1 = min(S_rows, r0 + CH)
      a0, a1 = max(r0 - 1, 0), min(r1 + 1, S_rows)
      F = filled[a0:a1]; Z = dem[a0:a1]; A = acc[a0:a1]; D = fd[a0:a1].long(); Wl = ws[a0:a1]; Lb = lab[a0:a1]
      o0, o1 = r0 - a0, r0 - a0 + (r1 - r0)                   # own rows inside the padded chunk
      Fo, Zo = F[o0:o1], Z[o0:o1]
      bad_fill += int((Fo < Zo).sum())
      inner = torch.zeros_like(Fo, dtype=torch.bool)
      rr = torch.arange(r0, r1, device=F.device).view(-1, 1)
      cc = torch.arange(S, device=F.device).view(1, -1)
      interior = (rr > 0) & (rr < S_rows - 1) & (cc > 0) & (cc < S - 1)
      bad_fill += int(((Fo != Zo) & ~interior).sum())
      upstream = torch.ones_like(A[o0:o1])
      wsdown = torch.zeros_like(Wl[o0:o1])
      moves = torch.zeros_like(interior)
      for q in range(8):
          # neighbour in direction q of every own cell (where it exists)
          rs, cs = DR[q], DC[q]
          ro0, ro1 = o0 + rs, o1 + rs
          src_r0, src_r1 = max(ro0, 0), min(ro1, F.shape[0])
          dst_r0 = src_r0 - ro0
          dst_r1 = dst_r0 + (src_r1 - src_r0)
          c_src0, c_src1 = max(cs, 0), S + min(cs, 0)
          c_dst0, c_dst1 = max(-cs, 0), S + min(-cs, 0)
          if src_r1 <= src_r0:
              continue
          nbF = F[src_r0:src_r1, c_src0:c_src1]
          own = (slice(dst_r0, dst_r1), slice(c_dst0, c_dst1))
          raised = (Fo[own] > Zo[own]) & interior[own]
          bad_fill += int((raised & (nbF < Fo[own])).sum())
          # accumulation: neighbour q flows into me iff its code is (q + 4) % 8
          into = D[src_r0:src_r1, c_src0:c_src1] == ((q + 4) & 7)
          upstream[own] += torch.where(into, A[src_r0:src_r1, c_src0:c_src1], torch.zeros_like(upstream[own]))
          # watersheds: my downstream cell is neighbour q iff my code is q
          mine = D[o0:o1][own] == q
          wsdown[own] = torch.where(mine, Wl[src_r0:src_r1, c_src0:c_src1], wsdown[own])
          moves[own] |= mine
      bad_acc += int((upstream != A[o0:o1]).sum())
      root_sum += float(A[o0:o1][~moves].sum())
      want = torch.where(Lb[o0:o1] != 0, Lb[o0:o1], torch.where(moves, wsdown, torch.zeros_like(wsdown)))
      bad_ws += int((want != Wl[o0:o1]).sum())
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
