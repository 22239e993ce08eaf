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
out = (ctypes.c_ulonglong * 16)()
dbg = raw.ms_nf_debug
dbg.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
import time
nflog = raw.ms_nf_log
nflog.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
for rep in range(3):
    dbg(None, 0, 1)
    nflog(None, None, 1)
    st = (ctypes.c_int64 * 8)()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rc = L.ms_fill_terrain_no_flats_dev(dem.data_ptr(), filled.data_ptr(), sh, dg, fnf.data_ptr(), S, S, st, sp)
    if rc != 0:
        print("rc", rc, L.ms_last_error()); break
    torch.cuda.synchronize(); t1 = time.perf_counter()
    dbg(out, 16, 0)
    v = max(st[1], 1)
    print("wall %.2f ms  visits %d  mean load %d cyc  mean solve %d cyc  mean iters %.2f  max solve %d cyc" %
          ((t1 - t0) * 1e3, st[1], out[1] // v, out[2] // v, out[3] / v, out[4]), " fp64-form visits", out[5], "reasons(warps F0/subnormal, W other binade, D range; tile range, weights, overflow)", list(out[6:12]))

log = np.zeros(4 * 262144, dtype=np.uint64)
nl = ctypes.c_uint(0)
nflog(log.ctypes.data_as(ctypes.c_void_p), ctypes.byref(nl), 0)
k = min(nl.value, 262144)
log = log[: 4 * k].reshape(k, 4)
os.makedirs("gpurun_out", exist_ok=True)
np.save("gpurun_out/nf_log_%d%s.npy" % (S, os.environ.get("MS_TAG", "")), log)
print("logged", k, "visits")
print("load phase: warp0 loads+convert %d cyc, reduce+sync %d cyc (means per visit)" % (out[13] // max(st[1], 1), out[14] // max(st[1], 1)))
print("sample out-of-range cell: W=%r F=%r count=%d" % (np.array([out[13]], dtype=np.uint64).view(np.float64)[0], np.array([out[14]], dtype=np.uint64).view(np.float64)[0], out[12]))
