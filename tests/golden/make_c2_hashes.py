#!/usr/bin/env python
"""Generate tests/golden/c2_hashes.json: SHA-256 of every raster and table the REFERENCE ITSELF produces on the
synthetic fractal DEM (malstroem_b200/synth.py, seed 1) at BASELINE configs[1] = 8192 x 8192 and at 2048 / 4096
(run in the build container only; baseline/_ref = the stock reference package, installed by baseline/install_ref.py,
with its own speedups.enable()).

Every value comes from the reference's stock functions (malstroem.algorithms.{fill,flow,label}, Cython path where the
reference has one, pure Python for label_max_index, label.py:135-166) with ONE exception: accumulated_flow re-traces
from every cell (29-157 s at 4096^2, hours at 8192^2), so above 2048^2 it is taken from the C port's
`fast=True` form, which tests/test_oracle_golden.py pins against the reference's own accumulation; at 2048^2 both are
run and asserted equal here.

Canonical bytes: C-contiguous array, float rasters with `+ 0.0` applied (folds -0.0 into +0.0, the one documented
deviation).  label_stats['sum'] is a float64 sum in raster order; the GPU's association differs,
so it is stored as a sample (every `SUM_STRIDE`-th label) + its total instead of a hash (tolerance 1e-6 relative).

    python tests/golden/make_c2_hashes.py [sizes...]        (default 2048 4096 8192; ~25 min, 12 GB)
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import toolchain  # noqa: E402
from malstroem_b200 import synth  # noqa: E402
from oracle import port  # noqa: E402

SUM_STRIDE = 997
OUT = os.path.join(HERE, "c2_hashes.json")


def sha(a):
    a = np.ascontiguousarray(a)
    if a.dtype.kind == "f":
        a = a + 0.0
    return hashlib.sha256(a.tobytes()).hexdigest()


def one(size, seed=1):
    alg, _, _, _, _ = toolchain.import_reference()
    from malstroem.algorithms import fill, flow, label, speedups
    speedups.enable()
    assert speedups.enabled
    t0 = time.time()
    dem = synth.fractal_dem(size, size, seed=seed)
    h = {"size": size, "seed": seed, "dem": sha(dem)}

    def lap(what):
        print("  %5d  %-28s %7.1f s" % (size, what, time.time() - t0), flush=True)

    filled = fill.fill_terrain(dem)
    lap("fill_terrain")
    depths = filled - dem
    short, diag = fill.minimum_safe_short_and_diag(dem)
    fnf = fill.fill_terrain_no_flats(dem, short, diag)
    lap("fill_terrain_no_flats")
    fd = flow.terrain_flowdirection(fnf, edges_flow_outward=True)
    lap("terrain_flowdirection")
    acc = port.accumulated_flow(fd, fast=True)
    if size <= 2048:
        assert np.array_equal(acc, flow.accumulated_flow(fd))
    lap("accumulated_flow")
    lab, n = label.connected_components(depths)
    st = label.label_stats(depths, lab)
    lap("connected_components + stats")
    ws = lab.copy()
    flow.watersheds_from_labels(fd, ws, unassigned=0)
    lap("watersheds_from_labels")
    cnt = label.label_count(ws)
    mi = label.label_min_index(fnf, lab, n)
    lap("label_min_index")
    ma = label.label_max_index(acc, lab, n)
    lap("label_max_index (pure Python)")
    h.update({
        "short": float(short), "diag": float(diag), "nlabels": int(n),
        "filled": sha(filled), "depths": sha(depths), "fnf": sha(fnf), "flowdir": sha(fd), "accum": sha(acc),
        "labels": sha(lab), "wsheds": sha(ws),
        "st_min": sha(st["min"]), "st_max": sha(st["max"]), "st_count": sha(st["count"].astype(np.int64)),
        "st_sum_stride": SUM_STRIDE, "st_sum_sample": [float(v) for v in st["sum"][::SUM_STRIDE]],
        "st_sum_total": float(st["sum"].sum()),
        "ws_count": sha(np.asarray(cnt, dtype=np.int64)),
        "ppmin_value": sha(mi["value"]), "ppmin_row": sha(mi["row"].astype(np.int64)),
        "ppmin_col": sha(mi["col"].astype(np.int64)),
        "ppmax_value": sha(ma["value"]), "ppmax_row": sha(ma["row"].astype(np.int64)),
        "ppmax_col": sha(ma["col"].astype(np.int64)),
        "seconds": round(time.time() - t0, 1),
    })
    return h


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [2048, 4096, 8192]
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for s in sizes:
        res[str(s)] = one(s)
        with open(OUT, "w") as f:
            json.dump(res, f, indent=1, sort_keys=True)
        print("wrote", OUT, "for", s, flush=True)


if __name__ == "__main__":
    main()
