#!/usr/bin/env python
"""Generate tests/golden/net188.npz + net_small.npz from the REFERENCE ITSELF (build container only) for the
SURVEY.md §8(f) rows: net.next_downstream_label / pourpoint_network (malstroem/algorithms/net.py:142-192) and
Network.rain_event (malstroem/network.py:75-129).

Sources of truth: /root/reference/malstroem/algorithms/net.py and /root/reference/malstroem/network.py imported
and run here (pure Python, interpreter recorded in the fixture because Python's sum() changed in 3.12), the
reference's golden rasters and tests/data/{pourpoints,nodes}.json, and the known answers of
tests/test_raster_net.py:8-21.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

from malstroem.algorithms import net, speedups          # noqa: E402
from malstroem.network import Network                   # noqa: E402

speedups.disable()
EVENTS = [10, 30, 100, 12.5]


def rain_arrays(nodes, events):
    """Run the reference Network; returns parent/area/cap in node order and [ne, n] value arrays (NaN = absent/None)."""
    ids = [n['nodeid'] for n in nodes]
    index = {nid: k for k, nid in enumerate(ids)}
    parent = np.array([-1 if n['dstrnodeid'] is None else index.get(n['dstrnodeid'], -2) for n in nodes], np.int64)
    area = np.array([float(n['wshed_area']) for n in nodes])
    cap = np.array([float(n['bspot_vol']) for n in nodes])
    nw = Network()
    nw.add_nodes(nodes)
    out = {k: np.full((len(events), len(nodes)), np.nan) for k in ("rainv", "spillv", "v", "pctv")}
    present = np.zeros(len(nodes), bool)
    order = None
    for e, mm in enumerate(events):
        evs = nw.rain_event(mm)
        if order is None:      # the order in which the reference returns the nodes (its evaluation order)
            order = np.array([index[ev['nodeid']] for ev in evs], np.int64)
        for ev in evs:
            k = index[ev['nodeid']]
            present[k] = True
            for key in out:
                out[key][e, k] = np.nan if ev[key] is None else ev[key]
    return dict(parent=parent, area=area, cap=cap, present=present, order=order, **out)


# ------------------------------------------------------------------------------------------------ dtm188
z = np.load(os.path.join(HERE, "dtm188.npz"))
fd, lab = z["flowdir_noflats"], z["labelled"]
g = {"python": np.array(sys.version_info[:3])}
# tests/test_raster_net.py:9-12
pps = [(1, 18), (5, 129), (7, 27), (7, 109), (9, 120), (12, 2), (17, 47), (33, 204), (12, 163), (14, 108),
       (18, 1), (23, 77), (20, 158), (21, 230), (24, 128), (26, 114), (34, 225), (27, 32), (30, 28), (44, 60)]
known = [None, None, 7, None, 2, 11, 13, 26, None, 4, None, 10, 9, 21, 2, 5, 21, 7, 18, 33]
geoms = []
for pp, want in zip(pps, known):
    lbl, geom = net.next_downstream_label(fd, lab, pp, background_label=0, geometry=True)
    assert lbl == want
    geoms.append(np.array(geom, np.int64))
g["t20_cells"] = np.array(pps, np.int64)
g["t20_down"] = np.array([-1 if k is None else k for k in known], np.int64)
g["t20_path_len"] = np.array([len(x) for x in geoms], np.int64)
g["t20_paths"] = np.concatenate(geoms)
# the whole pour-point table of pourpoints.json through pourpoint_network, background 0 and None
ppj = json.load(open(os.path.join(REF, "tests", "data", "pourpoints.json")))["features"]
for tag, bg in (("bg0", 0), ("bgnone", None)):
    nodes = net.pourpoint_network(fd, lab, ppj, bg)
    g["pp_ids"] = np.array([n['id'] for n in nodes], np.int64)
    g["pp_cells"] = np.array([n['pix'] for n in nodes], np.int64)
    g["pp_down_" + tag] = np.array([-1 if n['downstream_id'] is None else n['downstream_id'] for n in nodes], np.int64)
# raw (unfiltered) bluespots with their min-index pour points, as BluespotTool + StreamTool would chain them
raw_cells = np.stack([z["raw_minidx_row"], z["raw_minidx_col"]], 1)
nodes = net.pourpoint_network(fd, z["raw_labels"], [tuple(x) for x in raw_cells], 0)
g["raw_down_bg0"] = np.array([-1 if n['downstream_id'] is None else n['downstream_id'] for n in nodes], np.int64)
# nodes.json (the reference's golden node table, junction nodes included) through Network.rain_event
nj = [f['properties'] for f in json.load(open(os.path.join(REF, "tests", "data", "nodes.json")))["features"]]
g["nodes_id"] = np.array([n['nodeid'] for n in nj], np.int64)
g["events"] = np.array(EVENTS, np.float64)
for k, v in rain_arrays(nj, EVENTS).items():
    g["nodes_" + k] = v
# nodes.json as a graph on cells (junction ids depend on dict order, docs/cli.rst:206-208: not stable): per node its
# type (0 pour point, 1 junction), its cell and the cell of its downstream node (-1, -1: none)
by_id = {n['nodeid']: n for n in nj}
g["nodes_graph"] = np.array(sorted(
    (0 if n['nodetype'] == 'pourpoint' else 1, n['cell_row'], n['cell_col'],
     by_id[n['dstrnodeid']]['cell_row'] if n['dstrnodeid'] is not None else -1,
     by_id[n['dstrnodeid']]['cell_col'] if n['dstrnodeid'] is not None else -1) for n in nj), np.int64)
# and the reference's geometric_pourpoint_network run here on the golden rasters must be that graph
gn = net.geometric_pourpoint_network(fd, lab, [tuple(c) for c in g["pp_cells"].tolist()], 0)
gi = {n['id']: n for n in gn}
mine = np.array(sorted((0 if n['nodetype'] == 'pourpoint' else 1, n['pix'][0], n['pix'][1],
                        gi[n['downstream_id']]['pix'][0] if n['downstream_id'] is not None else -1,
                        gi[n['downstream_id']]['pix'][1] if n['downstream_id'] is not None else -1) for n in gn), np.int64)
assert np.array_equal(mine, g["nodes_graph"]), "geometric_pourpoint_network here != nodes.json as a graph"
# the exact output of the reference under this interpreter (insertion-ordered dicts make the junction ids deterministic)
g["geo_nodes"] = np.array([(n['id'], -1 if n['downstream_id'] is None else n['downstream_id'],
                            0 if n['nodetype'] == 'pourpoint' else 1, n['pix'][0], n['pix'][1], len(n['geometry']))
                           for n in gn], np.int64)
g["geo_paths"] = np.array([c for n in gn for c in n['geometry']], np.int64).reshape(-1, 2)
np.savez_compressed(os.path.join(HERE, "net188.npz"), **g)

# ------------------------------------------------------------------------------------------- small cases
sc = np.load(os.path.join(HERE, "small_cases.npz"))

rng = np.random.default_rng(7)
s = {}
k_out = 0
prefixes = sorted({k.split("_")[0] for k in sc.files if "_" in k and k.split("_")[0][:1] == "c"})
for pre in prefixes:
    keys = [k for k in sc.files if k.startswith(pre + "_")]
    if pre + "_flowdir" not in keys or pre + "_labels" not in keys:
        continue
    fdc, labc = sc[pre + "_flowdir"], sc[pre + "_labels"]
    rows, cols = fdc.shape
    cells = [(int(r), int(c)) for r, c in zip(sc[pre + "_minidx_row"], sc[pre + "_minidx_col"]) if r >= 0]
    cells += [(int(rng.integers(rows)), int(rng.integers(cols))) for _ in range(6)]
    for tag, bg in (("bg0", 0), ("bgnone", None)):
        down, plen, paths = [], [], []
        for pp in cells:
            lbl, geom = net.next_downstream_label(fdc, labc, pp, bg, geometry=True)
            down.append(-1 if lbl is None else lbl)
            plen.append(len(geom))
            paths.extend(geom)
        s["%s_down_%s" % (pre, tag)] = np.array(down, np.int64)
        s["%s_plen_%s" % (pre, tag)] = np.array(plen, np.int64)
        s["%s_paths_%s" % (pre, tag)] = np.array(paths, np.int64).reshape(-1, 2)
    s[pre + "_cells"] = np.array(cells, np.int64)
    # a node table as StreamTool (streams.py:66-100) would build it for the raw bluespots of this case
    nl = int(sc[pre + "_nlabels"])
    ppc = [(int(r), int(c)) for r, c in zip(sc[pre + "_minidx_row"], sc[pre + "_minidx_col"])]
    if all(r >= 0 for r, _ in ppc):
        nodes = net.pourpoint_network(fdc, labc, ppc, 0)
        wc = np.zeros(nl + 1, np.int64)
        cnt = sc[pre + "_wshed_count"]
        wc[:cnt.size] = cnt
        table = [dict(nodeid=n['id'], dstrnodeid=n['downstream_id'], wshed_area=float(wc[n['id']]) * 0.16,
                      bspot_vol=float(sc[pre + "_stats_sum"][n['id']]) * 0.16) for n in nodes]
        for key, v in rain_arrays(table, EVENTS).items():
            s["%s_rain_%s" % (pre, key)] = v
    k_out += 1
s["prefixes"] = np.array(prefixes)
# synthetic forests that stress the summation order: wide fan-in, long chains, dangling ids, zero capacities
for t in range(4):
    n = [1, 50, 400, 3000][t]
    parent = np.array([-1 if (i == 0 or rng.random() < 0.02) else int(rng.integers(0, i)) for i in range(n)])
    if t == 3:
        parent[1:200] = 0                                   # a hub
        parent[rng.integers(1, n, 5)] = -2                  # ids that are not nodes
    perm = rng.permutation(n)                               # insertion order unrelated to the tree order
    inv = np.argsort(perm)
    nodes = []
    for k in range(n):
        i = perm[k]
        p = parent[i]
        nodes.append(dict(nodeid=int(1000 + i), dstrnodeid=None if p == -1 else (int(1000 + p) if p >= 0 else 999999),
                          wshed_area=float(rng.random() * 1e4), bspot_vol=float(rng.choice([0.0, rng.random() * 50]))))
    for key, v in rain_arrays(nodes, EVENTS).items():
        s["forest%d_%s" % (t, key)] = v
np.savez_compressed(os.path.join(HERE, "net_small.npz"), **s)
print("net188.npz, net_small.npz written;", k_out, "small cases,", sys.version)
