"""torchrun, NF_STATS build (MS_LIB): the P2P no-flats solve of a banded raster — per rank the stage time, the tail
visits (log), and the remote pushes (count, mean latency).  usage: band_nf_stats.py rows_per_band cols"""
import ctypes, os, sys
os.environ["MS_BAND_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from malstroem_b200 import bands, _lib
from malstroem_b200.pipeline import synth_fractal
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
S, C = int(sys.argv[1]), int(sys.argv[2])
raw = ctypes.CDLL(_lib.LIB_PATH)
p = bands.BandPipeline(world * S, C, bands.DistComm(), device=local)
synth_fractal(S, C, seed=1, row0=rank * S, col0=0, device=local, out=p.dem)
for _ in range(2):
    p.run()
p.timing = {}
nflog = raw.ms_nf_log; nflog.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
dbg = raw.ms_nf_debug; dbg.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
hist = raw.ms_nf_hist; hist.argtypes = [ctypes.c_void_p, ctypes.c_int]
nflog(None, None, 1); dbg(None, 0, 1); hist(None, 1)
p.run()
hh = np.zeros(1024, dtype=np.uint64); hist(hh.ctypes.data_as(ctypes.c_void_p), 0)
out = (ctypes.c_ulonglong * 16)(); dbg(out, 16, 0)
log = np.zeros(4 * 262144, dtype=np.uint64); nl = ctypes.c_uint(0)
nflog(log.ctypes.data_as(ctypes.c_void_p), ctypes.byref(nl), 0)
k = min(nl.value, 262144)
lg = log[: 4 * k].reshape(k, 4).astype(np.int64)
msg = "rank %d: nf_init %.2f solve %.2f verify %.2f ms; visits %d; tail visits %d" % (
    rank, p.timing.get("nf_init", 0), p.timing.get("nf_p2p_solve", 0), p.timing.get("nf_verify", 0),
    p.stats.get("noflat_tile_visits", -1), k)
if k:
    t0 = lg[:, 0].min()
    msg += " spanning %.2f ms (first at +%.2f ms of the log epoch)" % ((lg[:, 2].max() - t0) / 1e6, 0.0)
    rounds = (lg[:, 3] >> 32) & 0xff
    relax = ((lg[:, 3] >> 40) & 0xffffff) / 1e3
    msg += "; per tail visit: load %.1f relax %.1f flush %.1f us, rounds %.2f" % (
        ((lg[:, 1] - lg[:, 0]) / 1e3).mean(), relax.mean(), ((lg[:, 2] - lg[:, 1]) / 1e3 - relax).mean(), rounds.mean())
msg += "; remote pushes %d, mean %.1f us" % (out[0], out[1] / max(out[0], 1) / 1e3)
if k:
    # the tail log for tools/nf_chain.py --analyse (columns pop, loaded, end [ns], tile, rounds, relax [ns])
    os.makedirs("gpurun_out", exist_ok=True)
    np.save("gpurun_out/nf_tail_band_r%d_of%d.npy" % (rank, world),
            np.stack([lg[:, 0], lg[:, 1], lg[:, 2], lg[:, 3] & 0xffffffff, (lg[:, 3] >> 32) & 0xff, (lg[:, 3] >> 40) & 0xffffff], axis=1))
# activity per millisecond: visits started, mean number of tiles in flight (of %d CTAs)
hv = hh.astype(np.float64).reshape(512, 2)
per = 15        # 15 bins of 65.5 us ~ 1 ms
lines = []
for a in range(0, 512, per):
    n, busy = hv[a:a + per, 0].sum(), hv[a:a + per, 1].sum()
    if n:
        lines.append("%5.1f ms: %6d visits, %5.0f in flight" % (a * 0.0655, n, busy / (per * 65536.0)))
msg += "\n  " + "\n  ".join(lines)
for r in range(world):
    dist.barrier()
    if r == rank:
        print(msg, flush=True)
dist.destroy_process_group()
