"""torchrun (any world size): BandPipeline.run_host — every rank's host rasters against a single-GPU run on rank 0,
rank 0's host tables against the device tables, and the time per call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from malstroem_b200 import bands
from malstroem_b200.pipeline import RasterPipeline, synth_fractal

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
R, C = world * S, S
p = bands.BandPipeline(R, C, bands.DistComm(), device=local)
dem = synth_fractal(p.rows, C, seed=1, row0=p.r0, col0=0, device=local).cpu()
for k in range(4):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    h = p.run_host(dem)
    dist.barrier(); dt = time.perf_counter() - t0
    if rank == 0:
        print("run_host %d: %.1f ms  (%.2f Gcell/s through host buffers)" % (k, dt * 1e3, R * C / dt / 1e9), flush=True)
ok = True
for name in ("filled", "depths", "flowdir", "accum", "labels", "wsheds"):
    ok &= bool(torch.equal(h[name], p.out[name].cpu()))
if rank == 0:
    for k, v in p.tables.items():
        ok &= bool(torch.equal(h["tables"][k], v.cpu()))
    ref = RasterPipeline(R, C, device=local)
    synth_fractal(R, C, seed=1, device=local, out=ref.dem)
    ref.run()
    for name in ("filled", "flowdir", "accum", "labels", "wsheds"):
        ok &= bool(torch.equal(h[name], ref.out[name][p.r0:p.r1].cpu()))
    print("RESULT", "OK" if ok else "FAIL")
flag = torch.tensor([0 if ok else 1], device="cuda"); dist.all_reduce(flag)
if rank == 0:
    print("all ranks", "OK" if int(flag) == 0 else "FAIL")
dist.destroy_process_group()
