"""Plain driver for profiling K2: synthetic DEM -> fill -> no-flats fill, REPS times (default 3).  MS_LIB selects the library."""
import ctypes, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from malstroem_b200 import _lib
from malstroem_b200.pipeline import synth_fractal

S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
L = _lib.lib(); L.ms_init(0)
dev = torch.device("cuda", 0)
dem = synth_fractal(S, S, seed=1)
filled = torch.empty_like(dem); depths = torch.empty_like(dem)
fnf = torch.empty((S, S), dtype=torch.float64, device=dev)
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
assert L.ms_fill_terrain_dev(dem.data_ptr(), filled.data_ptr(), depths.data_ptr(), S, S, sp) == 0
mv = np.float64(float(dem.abs().max())); sh = float((np.nextafter(mv, np.inf) - mv) * 1024); dg = sh * 2 ** 0.5
for rep in range(reps):
    st = (ctypes.c_int64 * 8)()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rc = L.ms_fill_terrain_no_flats_dev(dem.data_ptr(), filled.data_ptr(), sh, dg, fnf.data_ptr(), S, S, st, sp)
    assert rc == 0, L.ms_last_error()
    torch.cuda.synchronize()
    print("no-flats %.2f ms, visits %d" % ((time.perf_counter() - t0) * 1e3, st[1]))
print("checksum", float(fnf.sum()))
