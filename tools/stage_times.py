"""Wall-clock (host timer + sync) and kernel-event time per stage of the device-resident path."""
import ctypes, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from malstroem_b200 import _lib
from malstroem_b200.pipeline import synth_fractal

S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
L = _lib.lib(); L.ms_init(0)
dev = torch.device("cuda", 0)
dem = synth_fractal(S, S, seed=1)
if os.environ.get("MS_DEM") == "patho":          # BASELINE config 5 input (tools/c5_check.py)
    from c5_check import pathological_dem_device
    dem = pathological_dem_device(S)
n = S * S
filled = torch.empty((S, S), dtype=torch.float32, device=dev); depths = torch.empty_like(filled)
fnf = torch.empty((S, S), dtype=torch.float64, device=dev); fd = torch.empty((S, S), dtype=torch.uint8, device=dev)
acc = torch.empty((S, S), dtype=torch.float64, device=dev); lab = torch.empty((S, S), dtype=torch.int32, device=dev)
ws = torch.empty_like(lab)
cap = n // 4 + 2
tabs = [torch.empty(cap, dtype=torch.float64, device=dev) for _ in range(4)] + [torch.empty(cap, dtype=torch.int64, device=dev) for _ in range(4)]
nl = ctypes.c_int64(0)
st = torch.cuda.current_stream().cuda_stream
sp = ctypes.c_void_p(st)
mm = torch.empty(2, dtype=torch.float32, device=dev)

def prof():
    buf = ctypes.create_string_buffer(1 << 16); L.ms_profile_report(buf, len(buf))
    return sum(float(l.rsplit(" ", 3)[2]) for l in buf.value.decode().splitlines())

def timed(name, fn):
    torch.cuda.synchronize(); L.ms_profile(1)
    t0 = time.perf_counter(); rc = fn(); torch.cuda.synchronize(); t = time.perf_counter() - t0
    k = prof(); L.ms_profile(0)
    assert rc == 0, (name, L.ms_last_error())
    print("%-28s wall %8.2f ms   kernels %8.2f ms" % (name, t * 1e3, k))

for rep in range(reps):
    print("--- rep", rep)
    timed("fill", lambda: L.ms_fill_terrain_dev(dem.data_ptr(), filled.data_ptr(), depths.data_ptr(), S, S, sp))
    timed("minmax", lambda: L.ms_minmax_f32_dev(dem.data_ptr(), n, mm.data_ptr(), sp))
    lo, hi = mm.tolist(); import numpy as np
    mv = np.float64(max(abs(hi), abs(lo))); sh = float((np.nextafter(mv, np.inf) - mv) * 1024); dg = sh * 2 ** 0.5
    timed("noflats", lambda: L.ms_fill_terrain_no_flats_dev(dem.data_ptr(), filled.data_ptr(), sh, dg, fnf.data_ptr(), S, S, None, sp))
    timed("flowdir", lambda: L.ms_flowdir_dev(fnf.data_ptr(), fd.data_ptr(), S, S, 1, sp))
    timed("accum", lambda: L.ms_accumulated_flow_dev(fd.data_ptr(), acc.data_ptr(), S, S, sp))
    timed("cc", lambda: L.ms_connected_components_dev(depths.data_ptr(), 0, lab.data_ptr(), S, S, ctypes.byref(nl), sp))
    timed("label_stats", lambda: L.ms_label_stats_dev(depths.data_ptr(), 0, lab.data_ptr(), n, nl.value, tabs[0].data_ptr(), tabs[1].data_ptr(), tabs[2].data_ptr(), tabs[4].data_ptr(), sp))
    ws.copy_(lab)
    timed("watersheds", lambda: L.ms_watersheds_from_labels_dev(fd.data_ptr(), ws.data_ptr(), 4, S, S, 0, sp))
    timed("label_count", lambda: L.ms_label_count_dev(ws.data_ptr(), n, nl.value + 1, tabs[5].data_ptr(), sp))
    timed("min_index", lambda: L.ms_label_extreme_index_dev(fnf.data_ptr(), lab.data_ptr(), S, S, nl.value, 0, tabs[3].data_ptr(), tabs[6].data_ptr(), tabs[7].data_ptr(), sp))
    timed("max_index", lambda: L.ms_label_extreme_index_dev(acc.data_ptr(), lab.data_ptr(), S, S, nl.value, 1, tabs[3].data_ptr(), tabs[6].data_ptr(), tabs[7].data_ptr(), sp))
