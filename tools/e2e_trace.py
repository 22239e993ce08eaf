"""MS_SHIP_TRACE=1 python tools/e2e_trace.py [size]: when every raster of a host-buffer run (RasterPipeline.run_host) is
ready, when its device-to-host copy starts and how fast it runs (csrc/pipeline.cu ship())."""
import sys, time, torch
sys.path.insert(0, '.')
from malstroem_b200 import pipeline
n = int(sys.argv[1])
p = pipeline.RasterPipeline(n, n)
pipeline.synth_fractal(n, n, 1, out=p.dem)
h = p.host_buffers(); h["dem"].copy_(p.dem); torch.cuda.synchronize()
for k in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); p.run_host(); print("step %.1f ms" % ((time.perf_counter() - t0) * 1e3), file=sys.stderr)
