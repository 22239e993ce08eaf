"""torchrun: per-stage wall time of the banded run (MS_BAND_TIMING=1), rank 0 prints."""
import os, sys
os.environ["MS_BAND_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from malstroem_b200 import bands
from malstroem_b200.pipeline import synth_fractal
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
C = int(sys.argv[2]) if len(sys.argv) > 2 else S
p = bands.BandPipeline(world * S, C, bands.DistComm(), device=local)
synth_fractal(S, C, seed=1, row0=rank * S, col0=0, device=local, out=p.dem)
for _ in range(2):
    p.run()
p.timing = {}
reps = 2
for _ in range(reps):
    p.run()
if rank == 0:
    tot = 0
    for k, v in p.timing.items():
        print("%-12s %7.2f ms" % (k, v / reps)); tot += v / reps
    print("total        %7.2f ms" % tot, p.stats)
# the same run without the per-stage synchronisations: wall clock per step
import time
os.environ["MS_BAND_TIMING"] = "0"
torch.cuda.synchronize(); dist.barrier()
for mode in ("plain", "profiled"):
    from malstroem_b200 import _lib
    if mode == "profiled":
        _lib.lib().ms_profile(20000)
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    for _ in range(3):
        p.run()
    torch.cuda.synchronize(); dist.barrier(); dt = (time.perf_counter() - t0) / 3
    _lib.lib().ms_profile(0)
    if rank == 0:
        print("%s: %.2f ms per step without per-stage synchronisation" % (mode, dt * 1e3))
dist.destroy_process_group()
