import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np
from malstroem_b200 import _lib
from malstroem_b200.algorithms import fill, flow, label
from malstroem_b200.pipeline import synth_fractal
S = 8192
dem = synth_fractal(S, S, seed=1).cpu().numpy()
def T(name, f):
    t0 = time.perf_counter(); r = f(); print("%-34s %8.1f ms" % (name, (time.perf_counter() - t0) * 1e3)); return r
for rep in range(2):
    print("--- pass", rep); _lib.cache_clear()
    filled = T("fill_terrain", lambda: fill.fill_terrain(dem))
    depths = T("filled - dem (numpy)", lambda: filled - dem)
    del filled
    sd = T("minimum_safe_short_and_diag", lambda: fill.minimum_safe_short_and_diag(dem))
    fnf = T("fill_terrain_no_flats", lambda: fill.fill_terrain_no_flats(dem, *sd))
    fd = T("terrain_flowdirection", lambda: flow.terrain_flowdirection(fnf))
    del fnf
    acc = T("accumulated_flow", lambda: flow.accumulated_flow(fd))
    raw, n = T("connected_components", lambda: label.connected_components(depths))
    st = T("label_stats", lambda: label.label_stats(depths, raw))
    keep = T("filter (python)", lambda: (st["max"] > 0.05).tolist())
    comps = T("keep_labels", lambda: label.keep_labels(raw, keep))
    del raw
    lab, n = T("connected_components(bool)", lambda: label.connected_components(comps))
    T("label_stats 2", lambda: label.label_stats(depths, lab))
    ws = T("np.copy(lab)", lambda: np.copy(lab))
    T("watersheds_from_labels", lambda: flow.watersheds_from_labels(fd, ws, 0))
    T("label_count", lambda: label.label_count(ws))
    T("label_max_index", lambda: label.label_max_index(acc, lab, n))
    T("minmax 2", lambda: fill.minimum_safe_short_and_diag(dem))
    fnf = T("fill_terrain_no_flats 2", lambda: fill.fill_terrain_no_flats(dem, *sd))
    T("label_min_index", lambda: label.label_min_index(fnf, lab, n))
    del fnf, fd, acc, lab, ws, comps, depths
