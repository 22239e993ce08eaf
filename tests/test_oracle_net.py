"""The CPU oracle's §8(f) functions against outputs of the reference itself (net188.npz / net_small.npz) and the
known answers of /root/reference/tests/test_raster_net.py:8-21."""
import numpy as np

import net_cases
from conftest import _load
from oracle import port


def test_next_downstream_label_and_pourpoint_network_golden(dtm188):
    net_cases.check_net188(port, _load("net188.npz"), dtm188)


def test_net_small_cases(small_cases):
    net_cases.check_net_small(port, _load("net_small.npz"), small_cases)


def test_rain_events_golden():
    z, zs = _load("net188.npz"), _load("net_small.npz")
    assert tuple(z["python"][:2]) >= (3, 12)         # fixtures carry Neumaier sums (CPython >= 3.12 sum())
    net_cases.check_rain(port.rain_events, z, zs, exact=True)


def test_rain_plain_sum_within_tolerance():
    """sum_mode 0 (sum() before CPython 3.12) differs from the fixtures only in the last bits."""
    z, zs = _load("net188.npz"), _load("net_small.npz")
    net_cases.check_rain(lambda p, a, c, mm, _m: port.rain_events(p, a, c, mm, 0), z, zs, exact=False)


def test_junction_untangling_on_oracle_paths(dtm188):
    """The host half of geometric_pourpoint_network (malstroem_b200.algorithms.net.untangle_network, a restatement of
    net.py:43-139, 213-224) on paths from the CPU oracle == the reference's own output, junction ids and order
    included, and == the golden nodes.json as a graph."""
    from malstroem_b200.algorithms import net
    z = _load("net188.npz")
    fd, lab = dtm188["flowdir_noflats"], dtm188["labelled"]
    nodes = []
    for k, c in enumerate(z["pp_cells"].tolist()):
        lbl, geom = port.next_downstream_label(fd, lab, tuple(c), 0, geometry=True)
        nodes.append(dict(id=k, downstream_id=lbl, nodetype="pourpoint", pix=tuple(c), geometry=geom))
    out = net.untangle_network(nodes, int(lab.max()) + 1)
    net_cases.check_untangle_exact(out, z)
    assert len(out) == 117 and sum(n["nodetype"] == "junction" for n in out) == 12      # tests/test_raster_net.py:27-30


def test_network_evaluation_order_matches_reference():
    """Network.rain_event returns the nodes in the reference's evaluation order (network.py:100-129): the host-side
    order (no device involved) against the order recorded from the reference itself for nodes.json, the synthetic
    forests (hub, chains, unknown downstream ids) and the 40 small-case node tables."""
    from malstroem_b200 import network
    z, zs = _load("net188.npz"), _load("net_small.npz")

    def check(g, pre, ids=None):
        par = g[pre + "parent"]
        ids = np.arange(len(par)) if ids is None else ids
        nodes = [dict(nodeid=int(ids[k]), dstrnodeid=(None if par[k] == -1 else (int(ids[par[k]]) if par[k] >= 0 else 10 ** 9)),
                      wshed_area=1.0, bspot_vol=1.0) for k in range(len(par))]
        nw = network.Network()
        nw.add_nodes(nodes)
        assert nw._evaluation_order()[0] == g[pre + "order"].tolist(), pre

    check(z, "nodes_", z["nodes_id"])
    keys = ["forest%d_" % t for t in range(4)] + sorted(k[:-5] for k in zs if k.endswith("_rain_order"))
    assert len(keys) >= 44
    for pre in keys:
        check(zs, pre)
