"""Loader + driver loops for the reference's own compiled Cython modules (oracle/_ref/*.so).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The binaries are built from the sources under
/root/reference by oracle/build_ref.py; nothing here re-implements an algorithm except the trivial
Python driver loops that the reference keeps outside its native modules, which are restated (not
copied) with citations:

* fill_terrain / fill_terrain_no_flats: init + "UL, LR, UR, LL until one sweep changes nothing"
  driver  -> malstroem/algorithms/fill.py:102-109, :139-169, :207-232
* minimum_safe_short_and_diag -> fill.py:235-250
* terrain_flowdirection edge rule -> flow.py:118-139, :164-167
* watersheds_from_labels edge loop -> flow.py:398-412, _raster_utils.py:40-60
"""
import glob
import importlib.machinery
import importlib.util
import os
import sys
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")
_mods = {}


def available():
    return all(glob.glob(os.path.join(_REF, n + "*.so")) for n in ("_fill", "_flow", "_label"))


def _load(name):
    if name in _mods:
        return _mods[name]
    # _flow.pyx does `from ..dtypes import DTYPE_FLOWDIR` at import time: give it a stub package tree
    # (values are the reference's dtypes.py:20-31) unless a real malstroem package is already imported.
    if "malstroem" not in sys.modules:
        for pkg in ("malstroem", "malstroem.algorithms", "malstroem.algorithms.speedups"):
            m = types.ModuleType(pkg)
            m.__path__ = []
            sys.modules[pkg] = m
        d = types.ModuleType("malstroem.algorithms.dtypes")
        d.DTYPE_DTM = np.float32
        d.DTYPE_FILL = np.float32
        d.DTYPE_FILLNOFLAT = np.float64
        d.DTYPE_FLOWDIR = np.uint8
        d.DTYPE_ACCUM = np.float64
        sys.modules["malstroem.algorithms.dtypes"] = d
        sys.modules["malstroem.algorithms"].dtypes = d
    paths = glob.glob(os.path.join(_REF, name + "*.so"))
    if not paths:
        raise ImportError("oracle/_ref/%s*.so missing: run `python oracle/build_ref.py` where "
                          "/root/reference exists" % name)
    full = "malstroem.algorithms.speedups." + name
    loader = importlib.machinery.ExtensionFileLoader(full, paths[0])
    spec = importlib.util.spec_from_file_location(full, paths[0], loader=loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    _mods[name] = mod
    return mod


def _init_filled(dtm, dtype):
    # fill.py:102-109
    filled = np.full(dtm.shape, np.inf, dtype=dtype)
    filled[0, :] = dtm[0, :]
    filled[-1, :] = dtm[-1, :]
    filled[:, 0] = dtm[:, 0]
    filled[:, -1] = dtm[:, -1]
    return filled


def _sweep_until_quiet(sweep, rows, cols):
    # fill.py:139-169: four sweeps per round; the first sweep that reports no change ends everything
    maxrow, maxcol = rows - 2, cols - 2
    orders = ((1, maxrow, 1, maxcol), (maxrow, 1, maxcol, 1), (1, maxrow, maxcol, 1), (maxrow, 1, 1, maxcol))
    nsweeps = 0
    while True:
        for o in orders:
            nsweeps += 1
            if not sweep(*o):
                return nsweeps


def fill_terrain(dtm, return_sweeps=False):
    f = _load("_fill")
    filled = _init_filled(dtm, np.float32)
    n = _sweep_until_quiet(lambda a, b, c, d: f._fill_terrain(dtm, filled, a, b, c, d), *dtm.shape)
    return (filled, n) if return_sweeps else filled


def minimum_safe_short_and_diag(dem):
    # fill.py:235-250
    maxval = np.float64(max(abs(np.amax(dem)), abs(np.amin(dem))))
    nextval = np.nextafter(maxval, np.float64(np.inf))
    short = (nextval - maxval) * 1024
    diag = short * (2 ** 0.5)
    return short, diag


def fill_terrain_no_flats(dtm, short=0, diag=0):
    f = _load("_fill")
    filled = _init_filled(dtm, np.float64)
    _sweep_until_quiet(lambda a, b, c, d: f._fill_terrain_no_flats(dtm, filled, a, b, c, d, short, diag),
                       *dtm.shape)
    return filled


def terrain_flowdirection(terrain, edges_flow_outward=True):
    fl = np.asarray(_load("_flow").terrain_flow(terrain))
    if edges_flow_outward:
        # flow.py:118-139 (assignment order matters at the corners)
        fl[0, :] = 0
        fl[-1, :] = 4
        fl[:, 0] = 6
        fl[:, -1] = 2
        fl[0, 0] = 7
        fl[0, -1] = 1
        fl[-1, 0] = 5
        fl[-1, -1] = 3
    return fl


def accumulated_flow(flowdir):
    return np.asarray(_load("_flow").accumulated_flow(flowdir))


def watersheds_from_labels(flowdir, labelled, unassigned=0):
    f = _load("_flow")
    rows, cols = flowdir.shape
    # _raster_utils.py:40-60 edge order: (0,c),(maxr,c) per column, then (r,0),(r,maxc) per inner row
    for c in range(cols):
        f.assign_watersheds_upstream(flowdir, labelled, (0, c), unassigned)
        f.assign_watersheds_upstream(flowdir, labelled, (rows - 1, c), unassigned)
    for r in range(1, rows - 1):
        f.assign_watersheds_upstream(flowdir, labelled, (r, 0), unassigned)
        f.assign_watersheds_upstream(flowdir, labelled, (r, cols - 1), unassigned)


def label_stats(data, labelled, nlabels=None):
    return _load("_label").label_stats(data, labelled, nlabels)


def label_min_index(data, labelled, nlabels=None):
    return _load("_label").label_min_index(data, labelled, nlabels)
