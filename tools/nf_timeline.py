"""Timeline of the integer-raster no-flats solver (k_nf_solve_ir) from its per-visit log: needs a library built with
MS_BUILD_TAG=_st MS_NVCC_EXTRA=-DNF_STATS and MS_LIB pointing at it.  usage: python tools/nf_timeline.py [S] [bin_ms]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from malstroem_b200 import _lib
from malstroem_b200.pipeline import synth_fractal

S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
binms = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
C = int(sys.argv[3]) if len(sys.argv) > 3 else S          # columns (rows = S): a band-shaped raster
R0 = int(sys.argv[4]) if len(sys.argv) > 4 else 0         # first row of the window in the synthetic terrain
L = _lib.lib(); L.ms_init(0)
raw = ctypes.CDLL(_lib.LIB_PATH)
dev = torch.device("cuda", 0)
dem = synth_fractal(S, C, seed=1, row0=R0)
filled = torch.empty_like(dem); depths = torch.empty_like(dem)
fnf = torch.empty((S, C), dtype=torch.float64, device=dev)
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
assert L.ms_fill_terrain_dev(dem.data_ptr(), filled.data_ptr(), depths.data_ptr(), S, C, sp) == 0
mv = np.float64(float(dem.abs().max())); sh = float((np.nextafter(mv, np.inf) - mv) * 1024); dg = sh * 2 ** 0.5
nflog = raw.ms_nf_log
nflog.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
for rep in range(3):
    nflog(None, None, 1)
    st = (ctypes.c_int64 * 8)()
    import time as _t
    torch.cuda.synchronize(); _t0 = _t.perf_counter()
    rc = L.ms_fill_terrain_no_flats_dev(dem.data_ptr(), filled.data_ptr(), sh, dg, fnf.data_ptr(), S, C, st, sp)
    assert rc == 0, L.ms_last_error()
    torch.cuda.synchronize(); print('no-flats stage wall %.2f ms, visits %d' % ((_t.perf_counter() - _t0) * 1e3, st[1]))
log = np.zeros(4 * 262144, dtype=np.uint64)
nl = ctypes.c_uint(0)
nflog(log.ctypes.data_as(ctypes.c_void_p), ctypes.byref(nl), 0)
k = min(nl.value, 262144)
log = log[: 4 * k].reshape(k, 4).astype(np.int64)
t0 = log[:, 0].min()
pop, loaded, end = (log[:, 0] - t0) / 1e3, (log[:, 1] - t0) / 1e3, (log[:, 2] - t0) / 1e3      # us
rounds = (log[:, 3] >> 32) & 0xff
relax = ((log[:, 3] >> 40) & 0xffffff) / 1e3
print("visits logged %d (of %d), span %.2f ms" % (k, st[1], end.max() / 1e3))
print("per visit [us]: load mean %.2f  relax mean %.2f (p50 %.2f p90 %.2f max %.2f)  flush+queue mean %.2f  rounds mean %.2f"
      % ((loaded - pop).mean(), relax.mean(), np.median(relax), np.percentile(relax, 90), relax.max(),
         (end - loaded - relax).mean(), rounds.mean()))
print("window[ms]   running  started  load   relax  flush  rounds  relax/round")
edges = np.arange(0, end.max() / 1e3 + binms, binms)
for a in edges:
    b = a + binms
    m = (pop / 1e3 >= a) & (pop / 1e3 < b)
    if not m.any():
        continue
    mid = (a + b) / 2 * 1e3
    running = int(((pop <= mid) & (end > mid)).sum())
    r = rounds[m]
    print("%5.1f-%5.1f   %6d  %7d  %5.2f  %5.2f  %5.2f  %5.2f  %6.2f" % (a, b, running, m.sum(), (loaded - pop)[m].mean(), relax[m].mean(),
          (end - loaded - relax)[m].mean(), r.mean(), relax[m].sum() / max(r.sum(), 1)))
hist = np.bincount(rounds, minlength=12)
print("rounds histogram:", hist[:16].tolist())
# distribution of the visits by relax time and by rounds (share of visits / share of relax time)
tot_relax = relax.sum()
print("relax time [us]   visits   share   share of relax time")
for a, b in ((0, 2), (2, 5), (5, 10), (10, 20), (20, 40), (40, 80), (80, 1e9)):
    m = (relax >= a) & (relax < b)
    print("  %4g-%-6g %8d  %5.1f%%  %5.1f%%" % (a, b if b < 1e9 else float('inf'), m.sum(), 100.0 * m.mean(), 100.0 * relax[m].sum() / max(tot_relax, 1e-9)))
print("rounds   visits   mean relax us   share of relax time")
for r in range(0, 12):
    m = rounds == r
    if m.any():
        print("  %2d   %8d   %8.2f   %5.1f%%" % (r, m.sum(), relax[m].mean(), 100.0 * relax[m].sum() / max(tot_relax, 1e-9)))
first = np.zeros(k, dtype=bool)
seen = set()
order = np.argsort(pop)
tiles = (log[:, 3] & 0xffffffff)
for i in order:
    t = int(tiles[i])
    if t not in seen:
        seen.add(t)
        first[i] = True
print("first visits %d: mean relax %.2f us, rounds %.2f | re-visits %d: mean relax %.2f us, rounds %.2f, no-change (1 round) %.1f%%"
      % (first.sum(), relax[first].mean(), rounds[first].mean(), (~first).sum(), relax[~first].mean(), rounds[~first].mean(),
         100.0 * (rounds[~first] <= 1).mean()))
