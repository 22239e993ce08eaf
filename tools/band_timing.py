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
# per stage: wall time (device synchronised at the stage boundary) and the time inside OUR kernels (library profiler);
# the difference is collectives, torch glue kernels, host work and launch gaps
import ctypes, time as _time
from malstroem_b200 import _lib as _L
kern, perk = {}, {}
_orig_tick = p._tick


def _tick(name):
    _orig_tick(name)
    buf = ctypes.create_string_buffer(1 << 16)
    _L.lib().ms_profile_report(buf, len(buf))
    ms = 0.0
    for line in buf.value.decode().splitlines():
        kname, cnt, kms, _units = line.rsplit(" ", 3)
        ms += float(kms)
        d = perk.setdefault(name, {})
        d[kname] = d.get(kname, 0.0) + float(kms)
    kern[name] = kern.get(name, 0.0) + ms
    p._t_last = _time.perf_counter()


p._tick = _tick
_L.lib().ms_profile(20000)
for _ in range(reps):
    p.run()
_L.lib().ms_profile(0)
p._tick = _orig_tick
if rank == 0:
    tot = ktot = 0
    for k, v in p.timing.items():
        print("%-16s %7.2f ms   own kernels %7.2f ms   other %6.2f ms" % (k, v / reps, kern.get(k, 0) / reps, (v - kern.get(k, 0)) / reps))
        tot += v / reps
        ktot += kern.get(k, 0) / reps
        top = sorted(perk.get(k, {}).items(), key=lambda kv: -kv[1])[:7]
        if top:
            print("      " + "  ".join("%s %.2f" % (a, b / reps) for a, b in top))
    print("total            %7.2f ms   own kernels %7.2f ms   other %6.2f ms" % (tot, ktot, tot - ktot), p.stats)
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
