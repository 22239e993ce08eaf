"""ctypes front-end of oracle/ms_oracle.c — the CPU restatement ("port") of the hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Same numpy signatures as malstroem.algorithms
(fill.py, flow.py, label.py) so parity tests read like the reference's own tests.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "ms_oracle.c")
_OUT = os.path.join(_HERE, "_build", "libms_oracle.so")
_lib = None

c_i64 = ctypes.c_int64
c_p = ctypes.c_void_p


def build(force=False):
    if not force and os.path.exists(_OUT) and os.path.getmtime(_OUT) >= os.path.getmtime(_SRC):
        return _OUT
    os.makedirs(os.path.dirname(_OUT), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", _OUT, _SRC, "-lm"])
    return _OUT


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.orc_label.restype = c_i64
    return _lib


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def fill_terrain(dtm, return_sweeps=False):
    dtm = _c(dtm, np.float32)
    out = np.empty(dtm.shape, np.float32)
    n = lib().orc_fill_terrain(_p(dtm), _p(out), c_i64(dtm.shape[0]), c_i64(dtm.shape[1]))
    return (out, n) if return_sweeps else out


def minimum_safe_short_and_diag(dem):
    # fill.py:235-250
    maxval = np.float64(max(abs(np.amax(dem)), abs(np.amin(dem))))
    short = (np.nextafter(maxval, np.float64(np.inf)) - maxval) * 1024
    return short, short * (2 ** 0.5)


def fill_terrain_no_flats(dtm, short=0, diag=0):
    dtm = _c(dtm, np.float32)
    out = np.empty(dtm.shape, np.float64)
    lib().orc_fill_terrain_no_flats(_p(dtm), _p(out), c_i64(dtm.shape[0]), c_i64(dtm.shape[1]),
                                    ctypes.c_double(short), ctypes.c_double(diag))
    return out


def terrain_flowdirection(terrain, edges_flow_outward=True):
    t = _c(terrain, np.float64)
    out = np.empty(t.shape, np.uint8)
    lib().orc_flowdir(_p(t), _p(out), c_i64(t.shape[0]), c_i64(t.shape[1]), int(bool(edges_flow_outward)))
    return out


def accumulated_flow(flowdir, fast=False):
    fd = _c(flowdir, np.uint8)
    out = np.empty(fd.shape, np.float64)
    lib().orc_accum(_p(fd), _p(out), c_i64(fd.shape[0]), c_i64(fd.shape[1]), 1 if fast else 0)
    return out


def watersheds_from_labels(flowdir, labelled, unassigned=0):
    fd = _c(flowdir, np.uint8)
    lab = labelled.astype(np.int64)
    lib().orc_watersheds(_p(fd), _p(lab), c_i64(fd.shape[0]), c_i64(fd.shape[1]), c_i64(int(unassigned)))
    labelled[...] = lab.astype(labelled.dtype)


def connected_components(data):
    data = np.asarray(data)
    fg = np.ascontiguousarray(data != 0).view(np.uint8)
    out = np.empty(data.shape, np.int32)
    n = lib().orc_label(_p(fg), _p(out), c_i64(data.shape[0]), c_i64(data.shape[1]))
    return out, int(n)


STATS_DTYPE = [('min', np.float64), ('max', np.float64), ('sum', np.float64), ('count', np.int64)]
INDEX_DTYPE = [('value', np.float64), ('row', np.int64), ('col', np.int64)]


def label_stats(data, labelled, nlabels=None):
    if not nlabels:
        nlabels = int(np.max(labelled))
    d = _c(data, np.float64)
    lab = _c(labelled, np.int64)
    mn, mx, sm = (np.empty(nlabels + 1, np.float64) for _ in range(3))
    cnt = np.empty(nlabels + 1, np.int64)
    if lib().orc_label_stats(_p(d), _p(lab), c_i64(d.size), c_i64(nlabels), _p(mn), _p(mx), _p(sm), _p(cnt)):
        raise IndexError("label outside [0, nlabels]")
    out = np.zeros(nlabels + 1, dtype=STATS_DTYPE)
    out['min'], out['max'], out['sum'], out['count'] = mn, mx, sm, cnt
    return out


def _extreme(data, labelled, nlabels, want_max):
    if not nlabels:
        nlabels = int(np.max(labelled))
    d = _c(data, np.float64)
    lab = _c(labelled, np.int64)
    val = np.empty(nlabels + 1, np.float64)
    row = np.empty(nlabels + 1, np.int64)
    col = np.empty(nlabels + 1, np.int64)
    if lib().orc_label_extreme_index(_p(d), _p(lab), c_i64(d.shape[0]), c_i64(d.shape[1]), c_i64(nlabels),
                                     int(want_max), _p(val), _p(row), _p(col)):
        raise IndexError("label outside [0, nlabels]")
    out = np.zeros(nlabels + 1, dtype=INDEX_DTYPE)
    out['value'], out['row'], out['col'] = val, row, col
    return out


def label_min_index(data, labelled, nlabels=None):
    return _extreme(data, labelled, nlabels, False)


def label_max_index(data, labelled, nlabels=None):
    return _extreme(data, labelled, nlabels, True)


def label_count(labelled):
    lab = _c(labelled, np.int64).ravel()
    nb = int(lab.max()) + 1
    cnt = np.empty(nb, np.int64)
    if lib().orc_label_count(_p(lab), c_i64(lab.size), c_i64(nb), _p(cnt)):
        raise ValueError("negative label")
    return cnt


def keep_labels(labelled, keep_label, background=0):
    keep_label[background] = False      # label.py:94 mutates the caller's list
    keep = np.array(keep_label).astype(bool).view(np.uint8)
    lab = _c(labelled, np.int64)
    out = np.empty(lab.shape, np.uint8)
    if lib().orc_keep_labels(_p(lab.ravel()), c_i64(lab.size), _p(keep), c_i64(keep.size), _p(out)):
        raise IndexError("label outside keep list")
    return out.view(bool)


# ---------------------------------------------------------------------------------------- SURVEY §8(f)
def _pp_arrays(pour_points):
    # net.py:21-40 (_pourpoint_enumerator): json-type pour points or (row, col) pairs
    ids, cells = [], []
    for pid, pp in enumerate(pour_points):
        if isinstance(pp, dict) and 'properties' in pp:
            pid = pp['properties']['bspot_id']
            pp = (pp['properties']['cell_row'], pp['properties']['cell_col'])
        ids.append(pid)
        cells.append((int(pp[0]), int(pp[1])))
    cells = np.array(cells, dtype=np.int64).reshape(-1, 2)
    return ids, np.ascontiguousarray(cells[:, 0]), np.ascontiguousarray(cells[:, 1])


def _downstream(flowdir, labeled, rows_, cols_, background_label, geometry):
    fd = _c(flowdir, np.uint8)
    lab = _c(labeled, np.int64)
    n = rows_.size
    down = np.empty(n, np.int64)
    found = np.empty(n, np.uint8)
    plen = np.empty(n, np.int64)
    has_bg = background_label is not None
    args = (_p(fd), _p(lab), c_i64(fd.shape[0]), c_i64(fd.shape[1]), c_i64(n), _p(rows_), _p(cols_),
            c_i64(int(background_label) if has_bg else 0), int(has_bg), _p(down), _p(found), _p(plen))
    if lib().orc_next_downstream_labels(*args, None, None):
        raise RuntimeError("cyclic flow directions")
    paths = None
    if geometry:
        off = np.zeros(n + 1, np.int64)
        np.cumsum(plen, out=off[1:])
        cells = np.empty(int(off[-1]), np.int64)
        lib().orc_next_downstream_labels(*args, _p(cells), _p(off))
        paths = [[(int(i) // fd.shape[1], int(i) % fd.shape[1]) for i in cells[off[k]:off[k + 1]]] for k in range(n)]
    return down, found.astype(bool), paths


def next_downstream_label(flowdir, labeled, cell, background_label=None, geometry=False):
    # net.py:142-172
    r = np.array([int(cell[0])], np.int64)
    c = np.array([int(cell[1])], np.int64)
    down, found, paths = _downstream(flowdir, labeled, r, c, background_label, geometry)
    return (int(down[0]) if found[0] else None), (paths[0] if geometry else [])


def pourpoint_network(flowdir, labeled, pour_points, background_label=None):
    # net.py:175-192
    ids, r, c = _pp_arrays(pour_points)
    down, found, _ = _downstream(flowdir, labeled, r, c, background_label, False)
    return [dict(id=ids[k], downstream_id=(int(down[k]) if found[k] else None), nodetype='pourpoint',
                 pix=(int(r[k]), int(c[k]))) for k in range(len(ids))]


def rain_events(parent, area, cap, mm, sum_mode):
    """network.py:75-129 for all events at once.  parent: node index, -1 root, -2 unknown id.
    Returns dict of [ne, n] arrays (pctv NaN where the reference says None) and the `present` mask."""
    parent = _c(parent, np.int64)
    area = _c(area, np.float64)
    cap = _c(cap, np.float64)
    mm = _c(np.atleast_1d(mm), np.float64)
    n, ne = parent.size, mm.size
    out = {k: np.empty((ne, n), np.float64) for k in ("rainv", "spillv", "v", "pctv")}
    present = np.empty(n, np.uint8)
    if lib().orc_rain_events(c_i64(n), _p(parent), _p(area), _p(cap), c_i64(ne), _p(mm), int(sum_mode),
                             _p(out["rainv"]), _p(out["spillv"]), _p(out["v"]), _p(out["pctv"]), _p(present)):
        raise RuntimeError("cycle in the node network")
    out["present"] = present.astype(bool)
    return out
