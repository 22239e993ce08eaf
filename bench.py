#!/usr/bin/env python
"""bench.py — DEM cells/s end-to-end (fill -> D8 -> accum -> bluespot / watershed labels) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--size S] [--impl reference]

One "step" = one pass of the whole raster hot path (malstroem_b200.pipeline.RasterPipeline.run ->
ms_pipeline_dev) over one S x S synthetic fractal DEM per GPU (default S = 8192, BASELINE.json configs[1]).
`value` is timed with the DEM already resident in HBM; `e2e` is the same path through the host-buffer front end
(RasterPipeline.run_host: pinned host DEM -> H2D -> all stages -> D2H of every raster and table).
For N > 1 (torchrun, one rank per GPU) the ranks share ONE (N*S) x S raster split into N row bands of S rows
(malstroem_b200.bands.BandPipeline over NCCL: halo rows, the fill's boundary graph, the accumulation / watershed
exit forests, the bluespot boundary merge and the table all-reduces); per-GPU work is fixed, so the
scaling is weak.  `--independent` instead gives every rank its own S x S raster (no exchange).
`--impl reference` times the reference's own compiled Cython path (oracle/_ref) on the host, one core, on a
bounded window of the same DEM.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "DEM cells/sec end-to-end (fill->D8->accum->bluespot/watershed labels)"
UNIT = "Mcells/s"
BYTES_PER_CELL = 79.0     # compulsory traffic of the eight stages
# algorithmic bytes per cell of the stage each raster-wide kernel belongs to (SURVEY.md §8(d) / DESIGN.md §4)
STAGE_BYTES = {
    "k_descent_tile": 12, "k_forest_jump_list": 12, "k_ws_tile<L>": 9, "k_forest_jump": 12, "k_rootflag": 12, "k_catchment_ids": 12, "k_minedge<false>": 12, "k_minedge<true>": 12,
    "k_fill_final": 12, "k_scan_reduce<SELF>": 8, "k_scan_final<SELF>": 8,
    "k_nf_init": 12, "k_nf_seedcand": 12, "k_nf_solve<true>": 12, "k_nf_solve<false>": 12, "k_nf_solve_ir": 12, "k_nf_finish_ir": 12, "k_nf_init_tile": 12, "k_nf_verify": 12,
    "k_flowdir": 9, "k_acc_tile_a": 9, "k_acc_tile_c": 9, "k_acc_links": 9, "k_acc_node_trace": 9,
    "k_cc_tile<T>": 8, "k_cc_border": 8, "k_cc_flatten": 8, "k_cc_number": 8,
    "k_label_stats<T>": 8, "k_ws_ptr<L>": 9, "k_ws_assign<L>": 9, "k_label_count": 9,
    "k_extreme_key<true>": 12, "k_extreme_key<false>": 12, "k_extreme_index": 12, "k_minmax": 4,
    "k_tables_a<true>": 28, "k_tables_a<false>": 20, "k_tables_b<true>": 20, "k_tables_b<false>": 12,
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


# ------------------------------------------------------------------------------------ reference arm
def reference_stages(dem):
    """The reference's CPU path for the same stages (compiled Cython from oracle/_ref, scipy for the labelling
    exactly as malstroem/algorithms/label.py:35-39 does, and the C port for label_max_index, which the reference
    only has in pure Python, label.py:135-166)."""
    import scipy.ndimage
    from oracle import port, ref
    t0 = time.perf_counter()
    filled = ref.fill_terrain(dem)
    depths = filled - dem
    short, diag = ref.minimum_safe_short_and_diag(dem)
    fnf = ref.fill_terrain_no_flats(dem, short, diag)
    fd = ref.terrain_flowdirection(fnf, True)
    acc = ref.accumulated_flow(fd)
    lab, n = scipy.ndimage.label(depths, structure=np.ones((3, 3), int))
    lab = lab.astype(np.int32)
    st = ref.label_stats(depths, lab)
    ws = lab.copy()
    ref.watersheds_from_labels(fd, ws, 0)
    cnt = np.bincount(ws.ravel())
    mi = ref.label_min_index(fnf, lab, n)
    ma = port.label_max_index(acc, lab, n)
    return time.perf_counter() - t0, (filled, fnf, fd, acc, lab, ws, st, cnt, mi, ma)


def reference_available():
    from oracle import ref
    return ref.available()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from malstroem_b200 import synth
    s = args.ref_size
    dem = synth.fractal_dem(s, s, seed=1)
    if not reference_available():
        from oracle import port as _p   # noqa: F401  (fall back to the C port of the same algorithms)
        kind = "port"
    else:
        kind = "reference"
    times = []
    for i in range(args.warmup + args.steps):
        if kind == "reference":
            t, _ = reference_stages(dem)
        else:
            t = port_stages(dem)
        if i >= args.warmup:
            times.append(t)
    tot = sum(times)
    val = s * s * len(times) / tot / 1e6
    cores = 1
    sample = "%dx%d window (origin 0,0) of the seed-1 fractal DEM per step; reference is single-threaded" % (s, s)
    line = {"impl": "reference", "metric": METRIC, "value": round(val, 4), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * tot / len(times), 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic",
            "config": {"workload": "synthetic fractal DEM %dx%d float32: fill+depths, no-flats fill, D8, accum, "
                                   "bluespot labels, stats, watersheds, pour points" % (args.size, args.size),
                       "sample": sample},
            "cpu_baseline": {"value": round(val, 4), "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": round(val, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "host": {"cpus_available": len(os.sched_getaffinity(0))}}
    print(json.dumps(line))


def port_stages(dem):
    from oracle import port
    t0 = time.perf_counter()
    filled = port.fill_terrain(dem)
    depths = filled - dem
    short, diag = port.minimum_safe_short_and_diag(dem)
    fnf = port.fill_terrain_no_flats(dem, short, diag)
    fd = port.terrain_flowdirection(fnf, True)
    acc = port.accumulated_flow(fd)
    lab, n = port.connected_components(depths)
    port.label_stats(depths, lab)
    ws = lab.copy()
    port.watersheds_from_labels(fd, ws, 0)
    port.label_count(ws)
    port.label_min_index(fnf, lab, n)
    port.label_max_index(acc, lab, n)
    return time.perf_counter() - t0


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler(object):
    """SM clock and throttle reasons DURING the timed region, sampled through NVML from a background thread
    (an `nvidia-smi -lms` child process contends for the driver lock with the path's many small syncs and
    slowed the step several-fold; NVML reads are the same counters without that cost)."""

    def __init__(self, index, period=0.1):
        import threading
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        self.mode = os.environ.get("BENCH_CLOCKS", "nvml")
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None
            return
        if self.mode == "off":
            return
        self.period = period
        self._t = threading.Thread(target=self._loop, daemon=True)
        self._t.start()

    def _sample(self):
        nv = self.nv
        self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
            else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for name, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20),
                          ("hw_thermal_slowdown", 0x40), ("hw_power_brake_slowdown", 0x80)):
            if r & bit:
                self.reasons.add(name)

    def _loop(self):
        while not self._stop.is_set():
            try:
                self._sample()
            except Exception:
                pass
            self._stop.wait(self.period)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "source": "nvml thread, 100 ms"}
        if self.nv is None:
            out["source"] = "unavailable"
            return out
        if self._t is not None:
            self._stop.set()
            self._t.join(timeout=2)
        else:
            try:
                self._sample()
            except Exception:
                pass
        if self.samples:
            out["sm_mhz"] = float(np.median(self.samples))
            out["samples"] = len(self.samples)
        out["reasons"] = sorted(self.reasons)
        return out


# ---------------------------------------------------------------------------------------- our arm
def profile_report(lib):
    import ctypes
    buf = ctypes.create_string_buffer(1 << 16)
    lib.ms_profile_report(buf, len(buf))
    rows = {}
    for line in buf.value.decode().splitlines():
        name, n, ms, units = line.rsplit(" ", 3)
        rows[name] = (int(n), float(ms), int(units))
    return rows


def run_ours(args):
    import torch
    import torch.distributed as dist
    from malstroem_b200 import _lib
    from malstroem_b200.pipeline import RasterPipeline, synth_fractal

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (malstroem_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.lib()
    S = args.size
    n = S * S
    banded = world > 1 and not args.independent
    if banded:
        # one (world*S) x S raster, rank r owns rows [r*S, (r+1)*S)
        from malstroem_b200 import bands
        pipe = bands.BandPipeline(world * S, S, bands.DistComm(), device=local)
        assert pipe.rows == S and pipe.r0 == rank * S
    else:
        pipe = RasterPipeline(S, S, device=local, with_accum=True)
    # the band / the independent raster of rank r is the window of the synthetic terrain that starts at row r*S
    synth_fractal(S, S, seed=1, row0=rank * S, col0=0, device=local, out=pipe.dem)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    lib.ms_kernel_launches(1)
    for _ in range(args.warmup):
        pipe.run()
    barrier()
    per_step = int(lib.ms_kernel_launches(1)) // max(args.warmup, 1)
    # ---- timed region: K device-resident steps, CUDA events on the launching stream, per-kernel events on
    lib.ms_profile(max(2, int(per_step * args.steps * 1.5)))     # event pairs created before the clock starts
    lib.ms_kernel_launches(1)
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        pipe.run()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    launches = int(lib.ms_kernel_launches(1))
    prof = profile_report(lib)
    lib.ms_profile(0)
    # ---- e2e: host buffers in, host buffers out
    host = pipe.host_buffers()
    host["dem"].copy_(pipe.dem)
    torch.cuda.synchronize()
    pipe.run_host()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        pipe.run_host()
    barrier()
    e2e_s = time.perf_counter() - t0
    tms = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(tms[0]), float(tms[1])
    if rank == 0:
        peak, peak_kind = peaks()
        value = world * n * args.steps / (ms * 1e-3) / 1e6
        e2e = world * n * e2e_steps / (e2e_ms * 1e-3) / 1e6
        # dominant kernel: largest share of the kernel time inside the timed region
        tot_ms = sum(v[1] for v in prof.values())
        dom = max(prof.items(), key=lambda kv: kv[1][1])
        dname, (dn, dms, dunits) = dom
        units = dunits if dunits else dn * n
        bpc = STAGE_BYTES.get(dname, 8)
        achieved = bpc * units / (dms * 1e-3) / 1e9
        traffic = None
        try:        # per-launch DRAM bytes of this kernel at this size from the committed ncu capture, if there is one
            with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "top_kernel_traffic.json")) as f:
                ent = json.load(f).get(dname, {}).get(str(S))
            if ent and world == 1:
                traffic = {"dram_bytes_per_launch": ent["dram_bytes_per_launch"],
                           "vs_algorithmic": round(ent["dram_bytes_per_launch"] / (bpc * units / max(dn, 1)), 3),
                           "source": ent["source"]}
        except (OSError, ValueError):
            pass
        kernels = {k: {"launches": v[0], "ms_per_step": round(v[1] / args.steps, 3),
                       "share": round(v[1] / tot_ms, 4)} for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}
        line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic",
                "config": {"workload": "synthetic fractal DEM %dx%d float32 per GPU (seed 1, 1 mm quantised): "
                                       "fill+depths, no-flats fill, D8, accum, bluespot labels, stats, watersheds, "
                                       "pour points" % (S, S),
                           "parallelism": "1 GPU" if world == 1 else (
                               "%d row bands of %d rows of ONE %dx%d raster, one band per GPU, exchanges over NCCL"
                               % (world, S, world * S, S) if banded else "%d independent rasters, one per GPU" % world),
                           "l2": "inputs (%.0f MB DEM, %.1f GB working set) exceed the 126 MB L2; no flush" %
                                 (n * 4 / 1e6, n * 37 / 1e9),
                           "e2e_d2h": "filled f32, depths f32, flowdir u8, accum f64, bluespot labels i32, watersheds i32 "
                                      "(what DemTool / BluespotTool write, dem.py:67-93, bluespots.py:169,189) + all "
                                      "per-label tables; the float64 no-flats surface is an intermediate and stays in HBM",
                           "nlabels": pipe.nlabels, "stats": pipe.stats},
                "e2e": {"value": round(e2e, 2), "unit": UNIT, "h2d_bytes_per_step": pipe.bytes_h2d(),
                        "d2h_bytes_per_step": pipe.bytes_d2h(), "ms_per_step": round(e2e_ms / e2e_steps, 3)},
                "gpu_launches": launches,
                "roofline": {"bound": "hbm", "kernel": dname, "achieved": round(achieved, 2), "peak": peak,
                             "peak_kind": peak_kind, "unit": "GB/s", "frac": round(achieved / peak, 4),
                             "traffic": traffic, "bytes_per_unit": bpc, "units_per_launch": units // max(dn, 1),
                             "launch_ms": round(dms / max(dn, 1), 4), "share_of_kernel_time": round(dms / tot_ms, 4),
                             "pipeline_frac": round(value * 1e6 * BYTES_PER_CELL / (world * peak * 1e9), 5)},
                "kernels": kernels, "clocks": clocks}
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(args):
    from malstroem_b200 import synth
    s = args.ref_size_baseline
    dem = synth.fractal_dem(s, s, seed=1)
    if reference_available():
        t, _ = reference_stages(dem)
        kind = "reference"
    else:
        t = port_stages(dem)
        kind = "port"
    return {"value": round(s * s / t / 1e6, 4), "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "%dx%d window (origin 0,0) of the same seed-1 fractal DEM, all stages once (%.1f s)" % (s, s, t),
            "host_cpus_available": len(os.sched_getaffinity(0))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=8192)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--ref-size", type=int, default=1024, help="window edge per step of --impl reference")
    ap.add_argument("--ref-size-baseline", type=int, default=2048, help="window edge of the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--independent", action="store_true", help="N > 1: one independent raster per rank, no exchange")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl != "reference":
        args.warmup = 3
    # exactly one JSON line on stdout: libraries that write to fd 1 (NCCL's version banner) go to stderr instead
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
