"""Per-round statistics of the no-flats solver (needs a library built with MS_NVCC_EXTRA=-DNF_STATS)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from malstroem_b200 import _lib
from malstroem_b200.pipeline import synth_fractal

S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
L = _lib.lib(); L.ms_init(0)
raw = ctypes.CDLL(_lib.LIB_PATH) if hasattr(_lib, "LIB_PATH") else L
dev = torch.device("cuda", 0)
dem = synth_fractal(S, S, seed=1)
filled = torch.empty_like(dem); depths = torch.empty_like(dem)
fnf = torch.empty((S, S), dtype=torch.float64, device=dev)
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
assert L.ms_fill_terrain_dev(dem.data_ptr(), filled.data_ptr(), depths.data_ptr(), S, S, sp) == 0
mv = np.float64(float(dem.abs().max())); sh = float((np.nextafter(mv, np.inf) - mv) * 1024); dg = sh * 2 ** 0.5
out = (ctypes.c_ulonglong * (4 + 8 * 4096))()
dbg = raw.ms_nf_debug
dbg.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
for rep in range(2):
    dbg(None, 0, 1)
    st = (ctypes.c_int64 * 8)()
    assert L.ms_fill_terrain_no_flats_dev(dem.data_ptr(), filled.data_ptr(), sh, dg, fnf.data_ptr(), S, S, st, sp) == 0
    torch.cuda.synchronize()
    dbg(out, 4 + 8 * 4096, 0)
rounds = st[0]
print("rounds", rounds, "visits", st[1], "tile iterations", out[0], "block visits", out[1], "block iterations", out[2])
tot = 0
for r in range(rounds):
    n, ns, mload, msolve, mit, ssolve, sit = [out[4 + 8 * r + k] for k in range(7)]
    tot += ns
    print("round %3d tiles %6d  %8.1f us | max load %6d cyc  max solve %7d cyc  max iters %3d | mean solve %7d cyc  mean iters %.1f"
          % (r, n, ns / 1e3, mload, msolve, mit, ssolve // max(n, 1), sit / max(n, 1)))
print("sum of rounds %.2f ms" % (tot / 1e6))
