"""Per-kernel CUDA-event times of one device-resident step of the whole path at S x S (library profiler,
ms_profile / ms_profile_report), averaged over `reps` steps after warm-up.
usage: python tools/kernel_breakdown.py [S] [reps] [patho]"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from malstroem_b200 import _lib
from malstroem_b200.pipeline import RasterPipeline, synth_fractal

S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
L = _lib.lib()
p = RasterPipeline(S, S)
if len(sys.argv) > 3 and sys.argv[3] == "patho":
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from c5_check import pathological_dem_device
    p.dem.copy_(pathological_dem_device(S))
else:
    synth_fractal(S, S, seed=1, out=p.dem)
for _ in range(3):
    p.run()
torch.cuda.synchronize()
L.ms_kernel_launches(1)
p.run()
per = int(L.ms_kernel_launches(1))
L.ms_profile(int(per * reps * 1.5) + 16)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    p.run()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / reps * 1e3
buf = ctypes.create_string_buffer(1 << 16)
L.ms_profile_report(buf, len(buf))
L.ms_profile(0)
rows = []
for line in buf.value.decode().splitlines():
    name, n, ms, units = line.rsplit(" ", 3)
    rows.append((float(ms) / reps, int(n) // reps, name))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print("S=%d  wall %.2f ms/step (with per-kernel events)  kernels %.2f ms  %.2f Gcell/s  nlabels %d  stats %s"
      % (S, wall, tot, S * S / wall / 1e6, p.nlabels, p.stats))
for ms, n, name in rows:
    print("  %-28s %4d launches %9.3f ms  %5.1f%%" % (name, n, ms, 100 * ms / tot))
