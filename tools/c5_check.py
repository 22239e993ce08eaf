"""BASELINE config 5: pathological DEM (stepped plateaus = large exact flats, nested square craters centred on the
k*rows/8 band edges, a raster-wide flat; malstroem_b200/synth.py: pathological_dem, rebuilt here on the device with
torch because the numpy generator needs minutes and 30 GB at 32768^2) through the device-resident path, then the
bluespot network and the 10 / 30 / 100 mm rain events on device (SURVEY.md 8(f1,f2)).  Checks (test infrastructure):
  rasters      the certificates of tools/big_check.py (fill, accumulation, watersheds; the library's own no-flats stencil)
  network      downstream labels of a sample of pour points recomputed by the plain per-cell walker (< 64 start cells
               take that path) == the forest-based result for all pour points; no node is its own parent
  rain         per event: 0 <= v <= capacity, spill >= 0, spill > 0 only where the bluespot is full, and volume is
               conserved: sum(rain) == sum(v) + sum(spill of the roots), to 1e-9 relative
usage: python tools/c5_check.py [S] [reps]"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from malstroem_b200 import _lib
from malstroem_b200.pipeline import RasterPipeline, synth_fractal
from big_check import certify


def pathological_dem_device(S, seed=1, row0=0, nrows=None, device=0):
    """synth.pathological_dem(S, S, seed) with torch on the device (same integer-millimetre arithmetic); with
    row0 / nrows only the rows [row0, row0 + nrows) of it (one band of a row-band run)."""
    nrows = S if nrows is None else nrows
    dem = synth_fractal(nrows, S, seed=seed, row0=row0, device=device)
    out = torch.empty_like(dem)
    xs = torch.arange(S, device=dem.device, dtype=torch.int64).view(1, -1)
    CH = 2048
    for q0 in range(0, nrows, CH):
        q1 = min(nrows, q0 + CH)
        r0, r1 = q0, q1
        y = torch.arange(row0 + q0, row0 + q1, device=dem.device, dtype=torch.int64).view(-1, 1)
        mm = torch.round(dem[r0:r1].double() * 1000.0).long()
        mm = (mm // 5000) * 5000
        for k in range(1, 8):
            cy, cx = (k * S) // 8, ((2 * k + 1) * S) // 16
            rad = max(8, S // 24)
            d = torch.maximum((y - cy).abs(), (xs - cx).abs())
            ring = d * 8 // rad
            level = torch.where(ring % 2 == 0, 20000 + 3000 * ring, 60000 - 2000 * ring)
            mm = torch.where(d < rad, level, mm)
        strip = (y >= S // 3) & (y < S // 3 + max(2, S // 16))
        mm = torch.where(strip.expand_as(mm), torch.full_like(mm, 100000), mm)
        out[r0:r1] = mm.float() * np.float32(0.001)
    return out


if __name__ == "__main__":
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    p = RasterPipeline(S, S)
    p.dem.copy_(pathological_dem_device(S))
    if S <= 2048:       # the device generator against the numpy one
        from malstroem_b200 import synth
        assert np.array_equal(p.dem.cpu().numpy(), synth.pathological_dem(S, S, 1)), "device generator differs"
    torch.cuda.synchronize()
    events = [10.0, 30.0, 100.0]
    L = _lib.lib()
    for k in range(reps):
        hc = (ctypes.c_double * 4)(); L.ms_host_counters(hc, 1)
        if os.environ.get("MS_PROFILE"):
            L.ms_profile(1)
        t0 = time.perf_counter(); p.run(); torch.cuda.synchronize(); t1 = time.perf_counter()
        L.ms_host_counters(hc, 0)
        if os.environ.get("MS_PROFILE"):
            buf = ctypes.create_string_buffer(1 << 16); L.ms_profile_report(buf, len(buf)); L.ms_profile(0)
            rows_p = sorted((l.rsplit(" ", 3) for l in buf.value.decode().splitlines()), key=lambda r: -float(r[2]))
            print("   host: sync %.1f ms (%d), alloc %.1f ms (%d); kernels %.1f ms; top: %s" % (
                hc[0] * 1e3, hc[1], hc[2] * 1e3, hc[3], sum(float(r[2]) for r in rows_p),
                ", ".join("%s %.2f" % (r[0], float(r[2])) for r in rows_p[:5])))
        net = p.network(cell_area=0.16, events_mm=events); torch.cuda.synchronize(); t2 = time.perf_counter()
        print("run %d: rasters + tables %.1f ms (%.2f Gcell/s), network + %d rain events %.2f ms, nlabels %d, stats %s"
              % (k, (t1 - t0) * 1e3, S * S / (t1 - t0) / 1e9, len(events), (t2 - t1) * 1e3, p.nlabels, p.stats), flush=True)
    bad_fill, bad_acc, bad_ws, root_sum = certify(p)
    print("certificates: fill %d, accumulation %d (terminal sum %.0f vs N %d), watersheds %d violations"
          % (bad_fill, bad_acc, root_sum, S * S, bad_ws))
    ok = bad_fill == 0 and bad_acc == 0 and bad_ws == 0 and root_sum == S * S
    # ---- network: a sample through the plain walker
    n = p.nlabels + 1
    parent = net["parent"]
    ok &= bool((parent != torch.arange(n, device=parent.device, dtype=torch.int32)).all())
    ok &= bool(((parent >= -1) & (parent < n)).all())
    g = torch.Generator().manual_seed(5)
    idx = torch.randperm(n, generator=g)[:2000].to(parent.device)
    rows_, cols_ = p.table("ppmin_row")[idx].contiguous(), p.table("ppmin_col")[idx].contiguous()
    down = torch.empty(50, dtype=torch.int64, device=parent.device)
    found = torch.empty(50, dtype=torch.uint8, device=parent.device)
    sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    mism = 0
    for b in range(0, idx.numel(), 50):
        m = min(50, idx.numel() - b)
        _lib.check(_lib.lib().ms_pourpoint_network_dev(p.out["flowdir"].data_ptr(), p.out["labels"].data_ptr(), 4, S, S, m,
                                                       rows_[b:b + m].data_ptr(), cols_[b:b + m].data_ptr(), 0, 1,
                                                       down.data_ptr(), found.data_ptr(), sp), "pourpoint_network")
        want = torch.where(found[:m] != 0, down[:m], torch.full_like(down[:m], -1))
        mism += int((want != parent[idx[b:b + m]].long()).sum())
    print("network: %d nodes, %d roots, sample of %d pour points through the plain walker: %d mismatches"
          % (n, int((parent < 0).sum()), idx.numel(), mism))
    ok &= mism == 0
    # ---- rain: bounds, complementarity, conservation
    cap = p.table("st_sum") * 0.16
    roots = parent < 0
    for e, mm_ in enumerate(events):
        r, s, v = net["rainv"][e], net["spillv"][e], net["v"][e]
        bounds = int(((v < 0) | (v > cap) | (s < 0)).sum())
        compl = int(((s > 0) & (v < cap)).sum())
        lhs, rhs = float(r.sum()), float(v.sum() + s[roots].sum())
        rel = abs(lhs - rhs) / max(abs(lhs), 1e-300)
        print("rain %5.1f mm: rain volume %.6e, stored %.6e, leaving through the roots %.6e, conservation error %.2e, "
              "bound violations %d, spill-before-full %d, full bluespots %d"
              % (mm_, lhs, float(v.sum()), float(s[roots].sum()), rel, bounds, compl, int(((v >= cap) & (cap > 0)).sum())))
        ok &= bounds == 0 and compl == 0 and rel < 1e-9
    print("RESULT", "OK" if ok else "FAIL")
