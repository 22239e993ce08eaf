"""torchrun, one rank per GPU: ONE R x C synthetic fractal DEM as row bands (BASELINE configs 3 / 4, e.g. 65536^2 on
8 GPUs), timed, then certified per band (tools/big_check.certify on the band's rows, edge rows excluded) plus the
global identity  sum of acc over the raster border == R * C  (every path ends on a border cell that flows outward)."""
import importlib.util, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from malstroem_b200 import bands
from malstroem_b200.pipeline import synth_fractal
spec = importlib.util.spec_from_file_location("big_check", os.path.join(ROOT, "tools", "big_check.py"))
bc = importlib.util.module_from_spec(spec); spec.loader.exec_module(bc)

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
R = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
C = int(sys.argv[2]) if len(sys.argv) > 2 else R
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
p = bands.BandPipeline(R, C, bands.DistComm(), device=local)
synth_fractal(p.rows, C, seed=1, row0=p.r0, col0=0, device=local, out=p.dem)
torch.cuda.synchronize(); dist.barrier()
for k in range(reps):
    t0 = time.perf_counter(); p.run(); torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
    if rank == 0:
        print("run %d: %.1f ms  %.2f Gcell/s  nlabels %d  stats %s" % (k, dt * 1e3, R * C / dt / 1e9, p.nlabels, p.stats), flush=True)
# the bluespot network and 10 / 30 / 100 mm rain events on the replicated tables (SURVEY.md 8(f1,f2))
events = [10.0, 30.0, 100.0]
torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
net = p.network(cell_area=0.16, events_mm=events)
torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
if rank == 0:
    roots = net["parent"] < 0
    cap = p.tables["st_sum"] * 0.16
    msg = []
    for e, mm_ in enumerate(events):
        r, s_, v = net["rainv"][e], net["spillv"][e], net["v"][e]
        lhs, rhs = float(r.sum()), float(v.sum() + s_[roots].sum())
        viol = int(((v < 0) | (v > cap) | (s_ < 0) | ((s_ > 0) & (v < cap))).sum())
        msg.append("%g mm: conservation error %.1e, violations %d, full bluespots %d" % (mm_, abs(lhs - rhs) / max(lhs, 1e-300), viol, int(((v >= cap) & (cap > 0)).sum())))
    print("network + %d rain events: %.1f ms, %d nodes, %d roots; %s" % (len(events), dt * 1e3, net["parent"].numel(), int(roots.sum()), "; ".join(msg)), flush=True)
bad = torch.tensor(bc.certify(p, CH=1024, skip_top=1 if rank > 0 else 0, skip_bottom=1 if rank + 1 < world else 0)[:3],
                   dtype=torch.float64, device="cuda")
acc = p.out["accum"]
bsum = acc[:, 0].sum() + acc[:, -1].sum()
if rank == 0:
    bsum = bsum + acc[0, 1:-1].sum()
if rank + 1 == world:
    bsum = bsum + acc[-1, 1:-1].sum()
tot = torch.cat([bad, bsum.view(1)])
dist.all_reduce(tot)
if rank == 0:
    print("certificates over all bands: fill %d, accumulation %d, watersheds %d violations; border sum %.0f vs R*C %d"
          % (tot[0], tot[1], tot[2], tot[3], R * C))
    print("RESULT", "OK" if (tot[:3].sum() == 0 and tot[3] == R * C) else "FAIL")
dist.destroy_process_group()
