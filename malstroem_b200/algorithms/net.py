"""Mirror of malstroem/algorithms/net.py for the pour-point network: `next_downstream_label`
(net.py:142-172) and `pourpoint_network` (net.py:175-192) run on the device for all pour points at once;
`geometric_pourpoint_network` (net.py:195-224) takes every path from the device in one call and inserts the
junction nodes where paths merge on the host (net.py:43-139, list surgery over the paths)."""
import numpy as np

from .. import _lib


def _pourpoint_enumerator(pour_points):
    # net.py:21-40: json-type pour points carry their id and cell, anything else is a (row, col) pair
    for pid, pp in enumerate(pour_points):
        if isinstance(pp, dict) and 'properties' in pp:
            pid = pp['properties']['bspot_id']
            pp = (pp['properties']['cell_row'], pp['properties']['cell_col'])
        yield pid, pp


def _rasters(flowdir, labeled, what):
    fd = np.asarray(flowdir)
    lab = np.asarray(labeled)
    if fd.ndim != 2 or lab.shape != fd.shape:
        raise ValueError("%s: flowdir and labeled must be 2-D rasters of the same shape" % what)
    if fd.dtype != np.uint8:
        raise ValueError("%s: Buffer dtype mismatch, expected 'uint8' but got '%s'" % (what, fd.dtype))
    if lab.dtype == np.bool_ or not np.issubdtype(lab.dtype, np.integer):
        raise ValueError("%s: labeled must be an integer raster" % what)
    if lab.dtype not in (np.int32, np.int64):
        lab = lab.astype(np.int64 if lab.dtype.itemsize > 4 or lab.dtype == np.uint32 else np.int32)
    return np.ascontiguousarray(fd), np.ascontiguousarray(lab)


def _downstream(flowdir, labeled, cells, background_label, geometry, what):
    """cells: sequence of (row, col).  Returns (down int64[n], found bool[n], paths list or None)."""
    fd, lab = _rasters(flowdir, labeled, what)
    rows, cols = fd.shape
    n = len(cells)
    rc = np.array([(int(c[0]), int(c[1])) for c in cells], dtype=np.int64).reshape(n, 2)
    if n and (np.any(rc[:, 0] >= rows) or np.any(rc[:, 1] >= cols) or np.any(rc[:, 0] < -rows)
              or np.any(rc[:, 1] < -cols)):
        raise IndexError("%s: cell outside the raster" % what)      # labeled[cell[0], cell[1]], net.py:163
    # a negative index reads a label (numpy wrap-around) but trace_downstream yields nothing for it (flow.py:294):
    # such a cell answers None with an empty path, which is what the kernel returns for an out-of-raster start
    pr, pc = np.ascontiguousarray(rc[:, 0]), np.ascontiguousarray(rc[:, 1])
    down = np.zeros(n, np.int64)
    found = np.zeros(n, np.uint8)
    has_bg = background_label is not None
    L = _lib.lib()
    args = [_lib.ptr(fd), _lib.ptr(lab), lab.dtype.itemsize, rows, cols, n, _lib.ptr(pr), _lib.ptr(pc),
            int(background_label) if has_bg else 0, 1 if has_bg else 0, _lib.ptr(down), _lib.ptr(found)]
    paths = None
    with _lib.lock:
        if not geometry:
            _lib.check(L.ms_pourpoint_network(*args, None, None, 0), what)
        else:
            off = np.zeros(n + 1, np.int64)
            cap = max(1024, 64 * n)
            for _ in range(2):
                cells_out = np.empty(cap, np.int64)
                _lib.check(L.ms_pourpoint_network(*args, _lib.ptr(off), _lib.ptr(cells_out), cap), what)
                if off[n] <= cap:
                    break
                cap = int(off[n])
            r, c = np.divmod(cells_out[:off[n]], cols)
            flat = list(zip(r.tolist(), c.tolist()))
            paths = [flat[off[k]:off[k + 1]] for k in range(n)]
    return down, found.astype(bool), paths


def next_downstream_label(flowdir, labeled, cell, background_label=None, geometry=False):
    """net.next_downstream_label (net.py:142-172): (label or None, path of (row, col) cells — [] unless geometry)."""
    down, found, paths = _downstream(flowdir, labeled, [tuple(cell)], background_label, geometry,
                                     "next_downstream_label")
    return (int(down[0]) if found[0] else None), (paths[0] if geometry else [])


def pourpoint_network(flowdir, labeled, pour_points, background_label=None):
    """net.pourpoint_network (net.py:175-192): one node dict per pour point."""
    ids, cells = [], []
    for pid, pp in _pourpoint_enumerator(pour_points):
        ids.append(pid)
        cells.append(tuple(pp))
    down, found, _ = _downstream(flowdir, labeled, cells, background_label, False, "pourpoint_network")
    return [dict(id=ids[k], downstream_id=(int(down[k]) if found[k] else None), nodetype='pourpoint',
                 pix=tuple(cells[k])) for k in range(len(ids))]


def _common_flow_groups(nodes):
    """net.py:43-66 (min_common_cells = 2): nodes whose paths share their second-to-last cell flow together for at
    least the last two cells.  Groups in order of their first member; a path of two cells or fewer stands alone."""
    groups, rest = [], list(nodes)
    while rest:
        first = rest.pop(0)
        group = [first]
        if len(first['geometry']) > 2:
            key = first['geometry'][-2]
            for other in list(rest):
                if len(other['geometry']) > 2 and other['geometry'][-2] == key:
                    group.append(other)
                    rest.remove(other)
        groups.append(group)
    return groups


def _insert_junction(nodes, junction_id):
    """net.py:69-118: the cells all paths of `nodes` share at their downstream end become the path of a new junction
    node; the nodes are cut back to end at the junction's first cell and re-pointed to it."""
    paths = [list(n['geometry']) for n in nodes]
    shared = []
    while all(paths) and all(p[-1] == paths[0][-1] for p in paths):
        shared.append(paths[0][-1])
        for p in paths:
            p.pop()
    shared.reverse()                                     # back to flow order
    junction = dict(id=junction_id, downstream_id=nodes[0]['downstream_id'], nodetype='junction',
                    pix=tuple(shared[0]), geometry=shared)
    for n, p in zip(nodes, paths):
        n['downstream_id'] = junction_id
        n['geometry'] = p + [junction['pix']]
    return junction


def _untangle(nodes, next_id):
    """net.py:121-139: nodes flowing into one downstream node, with junction nodes inserted wherever their paths
    merge before they get there.  Yields (node, next free id), junctions before the nodes that flow into them."""
    for group in _common_flow_groups(nodes):
        if len(group) == 1:
            yield group[0], next_id
            continue
        junction = _insert_junction(group, next_id)
        next_id += 1
        yield junction, next_id
        for node, next_id in _untangle(group, next_id):
            yield node, next_id


def untangle_network(nodes, first_free_id):
    """The second half of net.geometric_pourpoint_network (net.py:213-224) on nodes that carry their `geometry`:
    grouped by downstream node in first-seen order, each group untangled, ids for junctions from `first_free_id`."""
    upstream = {}
    for node in nodes:
        upstream.setdefault(node['downstream_id'], []).append(node)
    final, next_id = [], int(first_free_id)
    for group in upstream.values():
        for node, next_id in _untangle(group, next_id):
            final.append(node)
    return final


def geometric_pourpoint_network(flowdir, labeled_bluespots, pour_points, background_label=None):
    """net.geometric_pourpoint_network (net.py:195-224).  Paths and downstream labels come from the device in one
    call; the junction nodes where paths merge are inserted on the host (list surgery over the paths, O(total path
    length))."""
    ids, cells = [], []
    for pid, pp in _pourpoint_enumerator(pour_points):
        ids.append(pid)
        cells.append(tuple(pp))
    down, found, paths = _downstream(flowdir, labeled_bluespots, cells, background_label, True,
                                     "geometric_pourpoint_network")
    nodes = [dict(id=ids[k], downstream_id=(int(down[k]) if found[k] else None), nodetype='pourpoint',
                  pix=tuple(cells[k]), geometry=paths[k]) for k in range(len(ids))]
    return untangle_network(nodes, int(np.max(labeled_bluespots) + 1))         # net.py:220
