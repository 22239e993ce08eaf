"""One device-resident step of the whole path at S x S after warm-up, bracketed by cudaProfilerStart/Stop
(ncu --profile-from-start off ... python tools/profile_step.py [S])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from malstroem_b200.pipeline import RasterPipeline, synth_fractal

S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
p = RasterPipeline(S, S)
synth_fractal(S, S, seed=1, out=p.dem)
for _ in range(3):
    p.run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
p.run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", p.nlabels, p.stats)
