"""Mirror of malstroem/network.py: `Network` with the reference's attributes and methods;
`rain_event` evaluates the whole forest on the device (csrc/network.cu), `rain_events` does several rain depths in
one pass.  `rain_events_arrays` is the array-level form the parity tests and malstroem_b200.pipeline use."""
import sys
from collections import defaultdict

import numpy as np

from . import _lib

# Python's sum() adds floats with Neumaier compensation from CPython 3.12 on; the reference calls sum()
# (network.py:86), so its results depend on the interpreter it runs under.  Match the running interpreter.
SUM_MODE = 1 if sys.version_info >= (3, 12) else 0


def rain_events_arrays(parent, wshed_area, bspot_vol, mm, sum_mode=None):
    """parent: node index of the downstream node, -1 root, -2 unknown id.  Returns dict of [n_events, n] float64
    arrays rainv / spillv / v / pctv (NaN = None) and the bool mask `present` (nodes the reference reaches)."""
    parent = np.ascontiguousarray(parent, dtype=np.int32)
    area = np.ascontiguousarray(wshed_area, dtype=np.float64)
    cap = np.ascontiguousarray(bspot_vol, dtype=np.float64)
    mm = np.ascontiguousarray(np.atleast_1d(mm), dtype=np.float64)
    n, ne = parent.size, mm.size
    if area.size != n or cap.size != n:
        raise ValueError("rain_events: parent, wshed_area and bspot_vol must have one entry per node")
    out = {k: np.full((ne, n), np.nan) for k in ("rainv", "spillv", "v", "pctv")}
    present = np.zeros(n, np.uint8)
    for e0 in range(0, ne, 16):                      # the library takes up to 16 events per pass
        e1 = min(ne, e0 + 16)
        part = {k: np.empty((e1 - e0, n)) for k in out}
        with _lib.lock:
            _lib.check(_lib.lib().ms_rain_events(n, _lib.ptr(parent), _lib.ptr(area), _lib.ptr(cap), e1 - e0,
                                                 _lib.ptr(mm[e0:e1].copy()), SUM_MODE if sum_mode is None else sum_mode,
                                                 _lib.ptr(part["rainv"]), _lib.ptr(part["spillv"]), _lib.ptr(part["v"]),
                                                 _lib.ptr(part["pctv"]), _lib.ptr(present)), "rain_events")
        for k in out:
            out[k][e0:e1] = part[k]
    out["present"] = present.astype(bool)
    return out


class Network(object):
    """Stream network (network.py:20-129): same attributes, `add_nodes`, `add_node`, `rain_event`."""

    def __init__(self):
        self.nodes = []
        self.nodes_index = {}
        self.root_nodes = []
        self.upstream_tree = defaultdict(list)
        self._node_rain_values = {}

    def add_nodes(self, nodes):
        for n in nodes:
            self.add_node(n)

    def add_node(self, node):
        # network.py:51-71
        self.nodes.append(node)
        node_id = node['nodeid']
        downstream_id = node['dstrnodeid']
        self.nodes_index[node_id] = node
        self.upstream_tree[downstream_id].append(node_id)
        if downstream_id is None:
            self.root_nodes.append(node_id)

    def _arrays(self):
        index = {n['nodeid']: k for k, n in enumerate(self.nodes)}
        if len(index) != len(self.nodes):
            raise ValueError("Network: duplicate node ids")
        parent = np.array([-1 if n['dstrnodeid'] is None else index.get(n['dstrnodeid'], -2) for n in self.nodes],
                          dtype=np.int32)
        area = np.array([float(n['wshed_area']) for n in self.nodes], dtype=np.float64)
        cap = np.array([float(n['bspot_vol']) for n in self.nodes], dtype=np.float64)
        return parent, area, cap

    def _evaluation_order(self):
        """Node indices in the order the reference evaluates (and returns) them, network.py:100-129: per root in
        insertion order, a stack walk upstream that visits the last-added upstream node first, evaluated in reverse."""
        index = {n['nodeid']: k for k, n in enumerate(self.nodes)}
        ups = {}
        for k, n in enumerate(self.nodes):
            ups.setdefault(n['dstrnodeid'], []).append(k)
        order = []
        for k, n in enumerate(self.nodes):
            if n['dstrnodeid'] is not None:
                continue
            walk, stack = [], [k]
            while stack:
                x = stack.pop()
                walk.append(x)
                stack.extend(ups.get(self.nodes[x]['nodeid'], ()))
            order.extend(reversed(walk))
        return order, index

    def rain_events(self, mmrains):
        """All events in one device pass: list (per event) of lists of event dicts as `rain_event` returns them."""
        mm = [float(m) for m in mmrains]
        out = rain_events_arrays(*self._arrays(), mm)
        order, _ = self._evaluation_order()
        assert int(out["present"].sum()) == len(order)
        res = []
        for e in range(len(mm)):
            r, s, v, p = (out[k][e] for k in ("rainv", "spillv", "v", "pctv"))
            # max(0, x) in the reference leaves the int 0 where nothing spills (network.py:92)
            res.append([dict(nodeid=self.nodes[k]['nodeid'], rainv=float(r[k]), spillv=float(s[k]) if s[k] else 0,
                             v=float(v[k]), pctv=None if np.isnan(p[k]) else float(p[k])) for k in order])
        return res

    def rain_event(self, mmrain):
        """network.py:113-129.  One dict per node reachable from a root (nodeid, rainv, spillv, v, pctv), in the
        reference's order (its evaluation order: leaves before the nodes they drain to, root by root)."""
        events = self.rain_events([mmrain])[0]
        self._node_rain_values = {e['nodeid']: e for e in events}
        return events
