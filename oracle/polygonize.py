"""TEST INFRASTRUCTURE — CPU oracle for SURVEY.md §8(f4), label polygonisation.

Only tests/, __graft_entry__.smoke() and bench.py's reference leg may import this; the product never does.

What it restates: /root/reference/malstroem/vector.py:42-87 calls `gdal.Polygonize(band, band.GetMaskBand(), layer, 0,
['8CONNECTED=8'])`.  GDAL (osgeo, any version: the reference pins none, requirements.txt / setup.py:72-79) is NOT in
/root/reference and not installed here, so this is a restatement of its PUBLISHED behaviour, not of code: one polygon
per 8-connected region of equal pixel value (masked pixels in none), exterior ring + one inner ring per hole, vertices
on pixel corners.  PARITY PIN: the reference's own known answer, tests/test_vector.py:18-20 — 113 features for
tests/data/labelled.tif (checked in tests/test_vector_cpu.py on the committed copy of that raster).  Unpinned (no GDAL
output to compare with): feature order, ring start vertex, winding.  The geometry itself is checked by rasterising the
rings back (`rasterize`), which must reproduce the raster exactly.

The method is deliberately not the device's: regions by breadth-first flood fill, rings by walking a per-region
dictionary of directed unit edges vertex by vertex, exterior / hole by the sign of the shoelace area.
"""
from collections import deque

import numpy as np


def regions(labels, connect8=True, nodata=None):
    """(region id raster, list of (value, first cell)) — ids in order of the first cell in raster order; -1 = masked."""
    rows, cols = labels.shape
    reg = np.full((rows, cols), -1, dtype=np.int64)
    info = []
    nb = [(-1, 0), (1, 0), (0, -1), (0, 1)]
    if connect8:
        nb += [(-1, -1), (-1, 1), (1, -1), (1, 1)]
    lab = labels.tolist()
    for r0 in range(rows):
        for c0 in range(cols):
            if reg[r0, c0] >= 0:
                continue
            v = lab[r0][c0]
            if nodata is not None and v == nodata:
                continue
            k = len(info)
            info.append((v, r0 * cols + c0))
            reg[r0, c0] = k
            q = deque([(r0, c0)])
            while q:
                r, c = q.popleft()
                for dr, dc in nb:
                    rr, cc = r + dr, c + dc
                    if 0 <= rr < rows and 0 <= cc < cols and reg[rr, cc] < 0 and lab[rr][cc] == v:
                        reg[rr, cc] = k
                        q.append((rr, cc))
    return reg, info


def _turn(d_in, d_out):
    """+1 right turn, -1 left turn, 0 straight (directions as (dr, dc) on the screen, rows growing downwards)"""
    cross = d_in[1] * d_out[0] - d_in[0] * d_out[1]      # x = col, y = row (y down): > 0 is clockwise on the screen
    return 1 if cross > 0 else (-1 if cross < 0 else 0)


def polygonize(labels, connect8=True, nodata=None):
    """[{'value', 'region' (first cell), 'rings': [exterior, hole, ...]}] — a ring is a list of (row, col) lattice
    corners, the region on the right-hand side, corners only, not closed; polygons ordered by first cell."""
    labels = np.asarray(labels)
    rows, cols = labels.shape
    reg, info = regions(labels, connect8, nodata)
    # directed boundary edges per region: start vertex -> list of end vertices
    out_edges = [dict() for _ in info]
    for r in range(rows):
        for c in range(cols):
            k = reg[r, c]
            if k < 0:
                continue
            def other(rr, cc):
                return not (0 <= rr < rows and 0 <= cc < cols and reg[rr, cc] == k)
            d = out_edges[k]
            if other(r - 1, c):
                d.setdefault((r, c), []).append((r, c + 1))
            if other(r, c + 1):
                d.setdefault((r, c + 1), []).append((r + 1, c + 1))
            if other(r + 1, c):
                d.setdefault((r + 1, c + 1), []).append((r + 1, c))
            if other(r, c - 1):
                d.setdefault((r + 1, c), []).append((r, c))
    polys = []
    for k, (v, first) in enumerate(info):
        d = out_edges[k]
        rings = []
        for start in sorted(d):
            while d.get(start):
                a = start
                b = d[a].pop(0)
                ring = [a]
                first_edge = (a, b)
                while True:
                    din = (b[0] - a[0], b[1] - a[1])
                    if b == first_edge[0]:
                        # back at the start vertex: the ring closes unless a saddle sends it on (never for the first
                        # vertex in sorted order: it is a convex corner of the ring)
                        break
                    cand = d[b]
                    if len(cand) == 1:
                        nxt = cand.pop(0)
                    else:
                        want = -1 if connect8 else 1       # saddle vertex: left turn joins the diagonal cells
                        turns = [_turn(din, (x[0] - b[0], x[1] - b[1])) for x in cand]
                        nxt = cand.pop(turns.index(want))
                    ring.append(b)
                    a, b = b, nxt
                rings.append(ring)
        # corners only
        simple = []
        for ring in rings:
            n = len(ring)
            keep = []
            for i in range(n):
                p, q, s = ring[i - 1], ring[i], ring[(i + 1) % n]
                if (q[0] - p[0], q[1] - p[1]) != (s[0] - q[0], s[1] - q[1]):
                    keep.append(q)
            simple.append(keep)
        ext = [g for g in simple if area2(g) > 0]
        holes = [g for g in simple if area2(g) < 0]
        assert len(ext) == 1, "region %d has %d exterior rings" % (k, len(ext))
        polys.append({"value": int(v), "region": int(first), "rings": ext + holes})
    return polys


def area2(ring):
    """twice the signed area; positive for a ring walked clockwise on the screen (x = col, y = row)"""
    s = 0
    n = len(ring)
    for i in range(n):
        (r0, c0), (r1, c1) = ring[i], ring[(i + 1) % n]
        s += c0 * r1 - c1 * r0
    return s


def canonical_ring(ring):
    """rotation-independent form: the lexicographically smallest rotation (a vertex may occur twice at saddles)"""
    ring = [tuple(int(x) for x in p) for p in ring]
    m = min(ring)
    best = None
    for i, p in enumerate(ring):
        if p == m:
            rot = tuple(ring[i:] + ring[:i])
            if best is None or rot < best:
                best = rot
    return best


def canonical(polys):
    """{(value, region): (exterior, sorted holes)} with every ring in canonical rotation"""
    out = {}
    for p in polys:
        rings = [canonical_ring(g) for g in p["rings"]]
        out[(p["value"], p["region"])] = (rings[0], tuple(sorted(rings[1:])))
    return out


def rasterize(polys, shape, background):
    """even-odd fill of every polygon's rings with its value; (raster, number of cells painted more than once)"""
    rows, cols = shape
    out = np.full(shape, background, dtype=np.int64)
    painted = np.zeros(shape, dtype=np.int32)
    for p in polys:
        tog = np.zeros((rows, cols + 1), dtype=np.uint8)
        for ring in p["rings"]:
            n = len(ring)
            for i in range(n):
                (r0, c0), (r1, c1) = ring[i], ring[(i + 1) % n]
                if c0 == c1 and r0 != r1:
                    tog[min(r0, r1):max(r0, r1), c0] ^= 1
        inside = np.bitwise_xor.accumulate(tog, axis=1)[:, :cols].astype(bool)
        out[inside] = p["value"]
        painted += inside
    return out, int((painted > 1).sum()), painted
