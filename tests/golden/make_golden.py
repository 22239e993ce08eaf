#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REFERENCE ITSELF (run in the build container only).

Sources of truth used here, none of which travel to the GPU box (hence the committed fixtures):
  * /root/reference/tests/data/*.tif, pourpoints.json  — the reference's own golden rasters
  * /root/reference/malstroem/algorithms (pure-Python path, speedups not built there) — semantics of record
  * oracle/_ref/*.so — the reference's compiled Cython modules (built by oracle/build_ref.py)
  * scipy.ndimage.label — the third-party labelling malstroem/algorithms/label.py:35-39 calls

Outputs:
  dtm188.npz        the 188x250 golden rasters + reference outputs the reference's tests only checksum
  small_cases.npz   ~40 small rasters (ties, flats, nested pits, NODIR cells, the -9999 plateau case of
                    tests/test_raster_fill.py:75-82) with pure-Python AND Cython reference outputs
  fractal256.npz    the synthetic fractal DEM (malstroem_b200/synth.py) 256x256 with Cython outputs
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

import cv2  # noqa: E402
import scipy.ndimage  # noqa: E402

from oracle import build_ref, ref as cy  # noqa: E402
from malstroem_b200 import synth  # noqa: E402

build_ref.build()
# the compiled modules first (they install stub packages), then the real pure-Python modules by path
cy._load("_fill"); cy._load("_flow"); cy._load("_label")
for k in [k for k in sys.modules if k.startswith("malstroem")]:
    if "speedups._" not in k:
        del sys.modules[k]
sys.path.insert(0, REF)
from malstroem.algorithms import fill as pyfill, flow as pyflow, label as pylabel, speedups  # noqa: E402
speedups.disable()          # the compiled modules are importable through the stubs: force the pure-Python path
assert not speedups.enabled and pyfill._fill_terrain.__module__ == "malstroem.algorithms.fill"


def rd(name):
    return cv2.imread(os.path.join(REF, "tests", "data", name), cv2.IMREAD_UNCHANGED)


def scipy_label(data):
    lab, n = scipy.ndimage.label(data, structure=np.ones((3, 3), int))
    return lab.astype(np.int32), n


def rec(a):
    return {k: np.asarray(a[k]) for k in a.dtype.names}


def all_outputs(dem, python_too):
    """Every hot-path function of the reference on one DEM."""
    o = {"dem": dem}
    o["filled_cy"] = cy.fill_terrain(dem)
    short, diag = cy.minimum_safe_short_and_diag(dem)
    o["short"], o["diag"] = np.float64(short), np.float64(diag)
    o["fnf_cy"] = cy.fill_terrain_no_flats(dem, short, diag)
    if python_too:
        o["filled_py"] = pyfill.fill_terrain(dem)
        s2, d2 = pyfill.minimum_safe_short_and_diag(dem)
        assert s2 == short and d2 == diag
        o["fnf_py"] = pyfill.fill_terrain_no_flats(dem, short, diag)
        fnf = o["fnf_py"]
        filled = o["filled_py"]
    else:
        fnf, filled = o["fnf_cy"], o["filled_cy"]
    o["flowdir"] = cy.terrain_flowdirection(fnf, True)
    o["flowdir_noedge"] = cy.terrain_flowdirection(fnf, False)
    o["accum"] = cy.accumulated_flow(o["flowdir"])
    if python_too:
        assert np.array_equal(o["accum"], pyflow.accumulated_flow(o["flowdir"]))
    depths = filled - dem
    o["depths"] = depths
    lab, n = scipy_label(depths)
    o["labels"], o["nlabels"] = lab, np.int64(n)
    st = cy.label_stats(depths, lab)
    for k, v in rec(st).items():
        o["stats_" + k] = v
    ws = lab.copy()
    cy.watersheds_from_labels(o["flowdir"], ws, 0)
    o["wsheds"] = ws
    if python_too:
        ws2 = lab.copy()
        pyflow.watersheds_from_labels(o["flowdir"], ws2, 0)
        assert np.array_equal(ws, ws2)
    o["wshed_count"] = pylabel.label_count(ws)
    mi = cy.label_min_index(fnf, lab, n)
    for k, v in rec(mi).items():
        o["minidx_" + k] = v
    if python_too or dem.size <= 70000:
        ma = pylabel.label_max_index(o["accum"], lab, n)
        for k, v in rec(ma).items():
            o["maxidx_" + k] = v
    return o


# ---------------------------------------------------------------------------------------- dtm188
g = {"dtm": rd("dtm.tif"), "filled": rd("filled.tif"), "depths": rd("depths.tif"),
     "filled_no_flats": rd("filled_no_flats.tif"), "flowdir_noflats": rd("flowdir_noflats.tif"),
     "labelled": rd("labelled.tif"), "wsheds": rd("wsheds.tif")}
pp = json.load(open(os.path.join(REF, "tests", "data", "pourpoints.json")))["features"]
pp = sorted(pp, key=lambda f: f["properties"]["bspot_id"])
for key in ("cell_row", "cell_col", "bspot_dmax", "bspot_area", "bspot_vol", "wshed_area"):
    g["pp_" + key] = np.array([f["properties"][key] for f in pp])
o = all_outputs(g["dtm"], python_too=False)
assert np.array_equal(o["filled_cy"], g["filled"]) and np.array_equal(o["fnf_cy"], g["filled_no_flats"])
assert np.array_equal(o["flowdir"], g["flowdir_noflats"]) and np.array_equal(o["depths"], g["depths"])
assert np.array_equal(np.asarray(pyfill.fill_terrain(g["dtm"])), g["filled"])
g["short"], g["diag"] = o["short"], o["diag"]
g["accum"] = o["accum"]
assert (g["accum"].min(), g["accum"].max(), g["accum"].sum()) == (1, 11158, 3578615)
# tests/test_raster_label.py:8-16: CC of (filled_no_flats - filled)
diff = g["filled_no_flats"] - g["filled"]
lab, n = scipy_label(diff)
assert n == 525 and (lab == 0).sum() == 40029 and lab.sum() == 1561377
g["cc_diff_labels"] = lab
# raw (unfiltered) bluespots of the depths raster and everything BluespotTool derives from them
for k in ("labels", "nlabels", "wsheds", "wshed_count", "stats_min", "stats_max", "stats_sum", "stats_count",
          "minidx_value", "minidx_row", "minidx_col", "maxidx_value", "maxidx_row", "maxidx_col"):
    g["raw_" + k] = o[k]
# the filtered golden labels (labelled.tif): stats / watersheds / pour points
ws = g["labelled"].copy(); cy.watersheds_from_labels(g["flowdir_noflats"], ws, 0)
assert np.array_equal(ws, g["wsheds"]) and ws.sum() == 2337891
st = cy.label_stats(g["depths"], g["labelled"])
for k, v in rec(st).items():
    g["lab_stats_" + k] = v
mi = cy.label_min_index(g["filled_no_flats"], g["labelled"])
assert np.array_equal(mi["row"], g["pp_cell_row"]) and np.array_equal(mi["col"], g["pp_cell_col"])
for k, v in rec(mi).items():
    g["lab_minidx_" + k] = v
ma = pylabel.label_max_index(g["accum"], g["labelled"])
for k, v in rec(ma).items():
    g["lab_maxidx_" + k] = v
g["lab_wshed_count"] = pylabel.label_count(ws)
np.savez_compressed(os.path.join(HERE, "dtm188.npz"), **g)

# ------------------------------------------------------------------------------------ small cases
rng = np.random.default_rng(20261018)
cases = []
for i in range(16):                                    # integer levels: ties, flats, nested pits
    r, c = int(rng.integers(4, 24)), int(rng.integers(4, 24))
    cases.append(rng.integers(0, int(rng.integers(2, 8)), (r, c)).astype(np.float32))
for i in range(8):                                     # rough continuous terrain
    r, c = int(rng.integers(8, 40)), int(rng.integers(8, 57))
    cases.append((rng.random((r, c)) * 10).astype(np.float32))
for i in range(6):                                     # crops of the synthetic fractal
    r, c = int(rng.integers(16, 48)), int(rng.integers(16, 64))
    cases.append(synth.fractal_dem(r, c, seed=i + 1, row0=1000 * i, col0=77 * i))
d = np.full((10, 10), -9999, np.float32); d[4:6, 4:6] = 0           # tests/test_raster_fill.py:75-82
cases.append(d)
d = np.zeros((12, 17), np.float32)                                  # one raster-wide flat
cases.append(d)
d = np.zeros((15, 15), np.float32); yy, xx = np.mgrid[:15, :15]     # concentric nested craters
d[:] = (np.maximum(abs(yy - 7), abs(xx - 7)) % 3).astype(np.float32)
cases.append(d)
d = np.full((9, 30), 5, np.float32); d[4, 1:29] = np.linspace(4, 1, 28).astype(np.float32)  # long channel to a pit
cases.append(d)
d = (np.add.outer(np.arange(20), np.arange(25)) * 0.25).astype(np.float32)   # tilted plane, no pits
cases.append(d)
d = synth.fractal_dem(4, 4, seed=9); cases.append(d)                # minimum size
d = synth.fractal_dem(4, 31, seed=9); cases.append(d)
d = synth.fractal_dem(33, 4, seed=9); cases.append(d)
d = (synth.fractal_dem(24, 24, seed=3) * np.float32(1e-38)).astype(np.float32)   # denormal depths
cases.append(d)
d = -synth.fractal_dem(20, 33, seed=4); cases.append(d)            # all negative

small = {"ncases": np.int64(len(cases))}
nquirk = 0
for i, dem in enumerate(cases):
    o = all_outputs(np.ascontiguousarray(dem), python_too=True)
    if not np.array_equal(o["filled_cy"], o["filled_py"]) or not np.array_equal(o["fnf_cy"], o["fnf_py"]):
        nquirk += 1                                  # Cython early exit; pure Python is the record
    for k, v in o.items():
        small["c%02d_%s" % (i, k)] = v
print("small cases:", len(cases), "cython-quirk cases:", nquirk)
# hand-made flow-direction rasters with NODIR cells / inward border cells for accumulation + watersheds
fds = []
for i in range(10):
    r, c = int(rng.integers(4, 20)), int(rng.integers(4, 20))
    if i < 6:
        dem = rng.integers(0, 6, (r, c)).astype(np.float64)      # flats -> interior NODIR cells
    else:
        dem = rng.random((r, c))                                 # no ties -> only pits are NODIR
    fd = cy.terrain_flowdirection(dem, bool(i % 2))
    lab, n = scipy_label(rng.random((r, c)) < 0.15)
    ws = lab.copy(); pyflow.watersheds_from_labels(fd, ws, 0)
    ws_cy = lab.copy(); cy.watersheds_from_labels(fd, ws_cy, 0)
    assert np.array_equal(ws, ws_cy)
    # accumulation over a NODIR cell is undefined in the reference (flow.py:230 raises TypeError,
    # _flow.pyx:203 reads an uninitialised delta), so it is only recorded where every cell flows
    acc = pyflow.accumulated_flow(fd) if (fd <= 7).all() else np.zeros((0, 0))
    small["f%02d_flowdir" % i] = fd
    small["f%02d_labels" % i] = lab
    small["f%02d_wsheds" % i] = ws
    small["f%02d_accum" % i] = acc
small["nflow"] = np.int64(10)
np.savez_compressed(os.path.join(HERE, "small_cases.npz"), **small)

# ------------------------------------------------------------------------------------- fractal256
dem = synth.fractal_dem(256, 256, seed=1)
o = all_outputs(dem, python_too=False)
assert np.array_equal(o["filled_cy"], np.asarray(pyfill.fill_terrain(dem)))   # converged (F2 check)
np.savez_compressed(os.path.join(HERE, "fractal256.npz"), **o)
for f in ("dtm188.npz", "small_cases.npz", "fractal256.npz"):
    print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
