"""Shared checks for the SURVEY.md §8(f) rows against tests/golden/net188.npz and net_small.npz (produced by the
reference itself, tests/golden/make_golden_net.py).  `impl` is either the CPU oracle (oracle.port) or the CUDA
path (malstroem_b200.algorithms.net + malstroem_b200.network): same function names, same arguments."""
import numpy as np

# interpreter that generated the fixtures: CPython >= 3.12 sums floats with Neumaier compensation
SUM_MODE_OF_FIXTURES = 1


def _none(x):
    return None if x < 0 else int(x)


def check_net188(net, z, dtm188):
    fd, lab = dtm188["flowdir_noflats"], dtm188["labelled"]
    # tests/test_raster_net.py:8-21 — the reference's own known answers, with the geometry
    off = np.concatenate([[0], np.cumsum(z["t20_path_len"])])
    for k, cell in enumerate(z["t20_cells"]):
        lbl, geom = net.next_downstream_label(fd, lab, tuple(cell), background_label=0, geometry=True)
        assert lbl == _none(z["t20_down"][k])
        assert geom and [tuple(map(int, c)) for c in geom] == [tuple(c) for c in z["t20_paths"][off[k]:off[k + 1]].tolist()]
        if lbl is not None:
            assert lab[geom[-1][0], geom[-1][1]] == lbl
        lbl2, geom2 = net.next_downstream_label(fd, lab, tuple(cell), background_label=0)
        assert lbl2 == lbl and geom2 == []
    cells = [tuple(c) for c in z["pp_cells"].tolist()]
    for tag, bg in (("bg0", 0), ("bgnone", None)):
        nodes = net.pourpoint_network(fd, lab, cells, bg)
        assert [n["id"] for n in nodes] == list(range(len(cells)))
        assert [n["downstream_id"] for n in nodes] == [_none(x) for x in z["pp_down_" + tag]]
        assert all(n["nodetype"] == "pourpoint" and n["pix"] == c for n, c in zip(nodes, cells))
    # json-type pour points (net.py:35-39)
    feats = [dict(properties=dict(bspot_id=int(i), cell_row=int(c[0]), cell_col=int(c[1])))
             for i, c in zip(z["pp_ids"], z["pp_cells"])]
    nodes = net.pourpoint_network(fd, lab, feats, 0)
    assert [n["id"] for n in nodes] == z["pp_ids"].tolist()
    assert [n["downstream_id"] for n in nodes] == [_none(x) for x in z["pp_down_bg0"]]
    raw_cells = list(zip(dtm188["raw_minidx_row"].tolist(), dtm188["raw_minidx_col"].tolist()))
    nodes = net.pourpoint_network(fd, dtm188["raw_labels"], raw_cells, 0)
    assert [n["downstream_id"] for n in nodes] == [_none(x) for x in z["raw_down_bg0"]]


def check_net_small(net, zs, small_cases):
    cases, _ = small_cases
    n = 0
    for i, case in enumerate(cases):
        pre = "c%02d" % i
        if pre + "_cells" not in zs:
            continue
        fd, lab = case["flowdir"], case["labels"]
        cells = [tuple(c) for c in zs[pre + "_cells"].tolist()]
        for tag, bg in (("bg0", 0), ("bgnone", None)):
            nodes = net.pourpoint_network(fd, lab, cells, bg)
            assert [n_["downstream_id"] for n_ in nodes] == [_none(x) for x in zs["%s_down_%s" % (pre, tag)]], (pre, tag)
            off = np.concatenate([[0], np.cumsum(zs["%s_plen_%s" % (pre, tag)])])
            for k in range(0, len(cells), 3):
                lbl, geom = net.next_downstream_label(fd, lab, cells[k], bg, geometry=True)
                assert lbl == _none(zs["%s_down_%s" % (pre, tag)][k])
                assert [list(map(int, c)) for c in geom] == zs["%s_paths_%s" % (pre, tag)][off[k]:off[k + 1]].tolist()
        n += 1
    assert n >= 30


def _check_rain(rain_events, g, pre, exact):
    out = rain_events(g[pre + "parent"], g[pre + "area"], g[pre + "cap"], g["events"] if "events" in g else EVENTS,
                      SUM_MODE_OF_FIXTURES)
    assert np.array_equal(out["present"], g[pre + "present"])
    m = g[pre + "present"]
    for k in ("rainv", "spillv", "v", "pctv"):
        a, b = out[k][:, m], g[pre + k][:, m]
        if exact:
            assert np.array_equal(a, b, equal_nan=True), (pre, k)
        else:
            np.testing.assert_allclose(a, b, rtol=1e-6, atol=1e-9, equal_nan=True)


EVENTS = np.array([10, 30, 100, 12.5])


def check_rain(rain_events, z, zs, exact=True):
    _check_rain(rain_events, z, "nodes_", exact)
    keys = sorted({k[:k.index("rain_") + 5] for k in zs if "_rain_" in k}) + ["forest%d_" % t for t in range(4)]
    assert len(keys) >= 30
    for pre in keys:
        _check_rain(rain_events, dict(zs, events=EVENTS), pre, exact)


def check_geometric_network(net, z, dtm188):
    """net.geometric_pourpoint_network (net.py:195-224) on the golden rasters == the reference's golden node table
    (tests/data/nodes.json: 105 pour points + 12 junction nodes, tests/test_raster_net.py:24-30), compared as a graph
    on cells because junction ids are not stable (docs/cli.rst:206-208)."""
    fd, lab = dtm188["flowdir_noflats"], dtm188["labelled"]
    nodes = net.geometric_pourpoint_network(fd, lab, [tuple(c) for c in z["pp_cells"].tolist()], 0)
    assert len(nodes) == 117 and sum(n["nodetype"] == "junction" for n in nodes) == 12
    by_id = {n["id"]: n for n in nodes}
    assert len(by_id) == len(nodes)
    graph = sorted((0 if n["nodetype"] == "pourpoint" else 1, int(n["pix"][0]), int(n["pix"][1]),
                    int(by_id[n["downstream_id"]]["pix"][0]) if n["downstream_id"] is not None else -1,
                    int(by_id[n["downstream_id"]]["pix"][1]) if n["downstream_id"] is not None else -1) for n in nodes)
    assert np.array_equal(np.array(graph, np.int64), z["nodes_graph"])
    for n in nodes:                                     # tests/test_raster_net.py:38-45
        assert tuple(n["geometry"][0]) == tuple(n["pix"])
        d = n["downstream_id"]
        if d is not None and by_id[d]["nodetype"] == "junction":
            assert tuple(n["geometry"][-1]) == tuple(by_id[d]["pix"])
    check_untangle_exact(nodes, z)


def check_untangle_exact(nodes, z):
    """A geometric network (list of node dicts in output order) == the reference's own output, ids and order included
    (fixture `geo_nodes` / `geo_paths`, generated under an insertion-ordered-dict interpreter)."""
    want, paths = z["geo_nodes"], z["geo_paths"]
    assert len(nodes) == len(want)
    off = 0
    for n, w in zip(nodes, want.tolist()):
        assert (n["id"], -1 if n["downstream_id"] is None else n["downstream_id"],
                0 if n["nodetype"] == "pourpoint" else 1, int(n["pix"][0]), int(n["pix"][1]), len(n["geometry"])) == tuple(w)
        assert [list(map(int, c)) for c in n["geometry"]] == paths[off:off + w[5]].tolist()
        off += w[5]
