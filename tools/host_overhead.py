"""Where does the host time of one device-resident step go?  (syncs, allocations, the rest)"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from malstroem_b200 import _lib
from malstroem_b200.pipeline import RasterPipeline, synth_fractal
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
L = _lib.lib()
p = RasterPipeline(S, S)
synth_fractal(S, S, seed=1, out=p.dem)
out = np.zeros(4)
for prof in (0, 0, 0, 1, 1, 0, 0):
    L.ms_profile(prof); L.ms_host_counters(None, 1); torch.cuda.synchronize()
    t0 = time.perf_counter(); p.run(); torch.cuda.synchronize(); t = time.perf_counter() - t0
    L.ms_host_counters(_lib.ptr(out), 1)
    k = 0.0
    if prof:
        buf = ctypes.create_string_buffer(1 << 16); L.ms_profile_report(buf, len(buf))
        k = sum(float(l.rsplit(" ", 3)[2]) for l in buf.value.decode().splitlines())
    print("prof=%d step %.1f ms | sync %.1f ms in %d syncs | alloc %.2f ms in %d allocs | kernels %.1f ms" %
          (prof, t * 1e3, out[0] * 1e3, out[1], out[2] * 1e3, out[3], k))
