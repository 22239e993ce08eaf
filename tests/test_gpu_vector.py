"""label polygonisation on the device (csrc/polygon.cu, malstroem_b200/vector.py) against the CPU
oracle (oracle/polygonize.py: flood fill + vertex-by-vertex walk + shoelace sign — none of the device's steps), the
reference's known answer (113 features, tests/test_vector.py:18-20) and, at sizes the oracle does not reach,
size-independent properties (one exterior ring per scipy region, ring areas = cell counts, closed rings)."""
import os

import numpy as np
import pytest
from scipy import ndimage

from malstroem_b200 import vector
from oracle import polygonize as P
from poly_cases import cases

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dtm188.npz")
TRANSFORM = (720000.0, 0.4, 0.0, 6193000.0, 0.0, -0.4)


def as_polys(rings):
    out = []
    for value, ks in rings.polygons():
        rr = []
        for k in ks:
            vr, vc = rings.ring_lattice(k)
            rr.append(list(zip(vr.tolist(), vc.tolist())))
        out.append({"value": value, "region": int(rings.region[ks[0]]), "rings": rr})
    return out


def check_against_oracle(a, connect8=True, nodata=None):
    rings = vector.polygonize_labels(a, connect8=connect8, nodata=nodata)
    ours = as_polys(rings)
    ref = P.polygonize(a, connect8=connect8, nodata=nodata)
    assert len(ours) == len(ref)
    assert [(p["value"], p["region"]) for p in ours] == [(p["value"], p["region"]) for p in ref]      # same order too
    assert P.canonical(ours) == P.canonical(ref)
    # what the device decides from the leader's side must be what the area says
    for p in ours:
        assert P.area2(p["rings"][0]) > 0 and all(P.area2(h) < 0 for h in p["rings"][1:])
    return rings


@pytest.mark.parametrize("name", sorted(cases()))
@pytest.mark.parametrize("connect8", [True, False])
def test_rings_equal_oracle(name, connect8):
    check_against_oracle(cases()[name], connect8)


def test_reference_fixture_113_features():
    lab = np.load(GOLDEN)["labelled"]
    rings = check_against_oracle(lab)
    feats = list(rings.features(TRANSFORM, "bspot_id"))
    assert len(feats) == 113
    assert sorted(set(f["properties"]["bspot_id"] for f in feats)) == list(range(105))
    for f in feats:
        assert f["type"] == "Feature" and f["geometry"]["type"] == "Polygon"
        for ring in f["geometry"]["coordinates"]:
            assert ring[0] == ring[-1] and len(ring) >= 5


def test_watersheds_fixture_and_nodata():
    ws = np.load(GOLDEN)["wsheds"]
    check_against_oracle(ws)
    rings = check_against_oracle(ws, nodata=0)
    assert 0 not in set(rings.value.tolist())
    check_against_oracle(cases()["random3"], connect8=False, nodata=2)
    everything = vector.polygonize_labels(np.zeros((5, 6), dtype=np.int32), nodata=0)
    assert len(everything) == 0 and list(everything.features()) == []


def test_other_integer_dtypes_and_errors():
    a = cases()["random1"]
    r32 = vector.polygonize_labels(a)
    for dt in (np.int64, np.uint8, np.int16):
        r = vector.polygonize_labels(a.astype(dt))
        assert np.array_equal(r.vrow, r32.vrow) and np.array_equal(r.vcol, r32.vcol) and np.array_equal(r.offset, r32.offset)
    with pytest.raises(ValueError):
        vector.polygonize_labels(a.astype(np.float32))
    with pytest.raises(ValueError):
        vector.polygonize_labels(a[0])


def _areas(rings):
    """signed twice-areas of all rings (vectorised shoelace)"""
    r, c = rings.vrow.astype(np.int64), rings.vcol.astype(np.int64)
    nxt = np.arange(len(r)) + 1
    last = rings.offset[1:] - 1
    nxt[last] = rings.offset[:-1]
    cross = c * r[nxt] - c[nxt] * r
    return np.add.reduceat(cross, rings.offset[:-1])


def test_pipeline_labels_2048_properties():
    """bluespot and watershed labels of a 2048^2 run: too big for the Python oracle; the properties pin the result"""
    import torch
    from malstroem_b200 import pipeline, synth
    n = 2048
    p = pipeline.RasterPipeline(n, n)
    p.dem.copy_(torch.from_numpy(synth.fractal_dem(n, n, seed=1)).cuda())
    p.run()
    for which in ("labels", "wsheds"):
        t = p.out[which]
        rings = vector.polygonize_labels_device(t)
        a = t.cpu().numpy()
        host = vector.polygonize_labels(a)
        for f in ("offset", "value", "cell", "region", "hole", "vrow", "vcol"):
            assert np.array_equal(getattr(rings, f), getattr(host, f)), f
        # one exterior ring per 8-connected region of equal value
        nreg = ndimage.label(a != 0, structure=np.ones((3, 3)))[1] if which == "labels" else None
        ext = rings.hole == 0
        if which == "labels":
            nzero = ndimage.label(a == 0, structure=np.ones((3, 3)))[1]
            assert int(ext.sum()) == nreg + nzero
        # ring areas: exterior minus holes = number of cells of the region
        area2 = _areas(rings)
        assert np.all(area2[ext] > 0) and np.all(area2[~ext] < 0)
        per_region = {}
        cells = np.zeros(int(rings.region.max()) + 1, dtype=np.int64)
        np.add.at(cells, rings.region, area2 // 2)
        assert int(cells.sum()) == n * n
        vals = rings.value[ext]
        count = np.bincount(a.ravel(), minlength=int(a.max()) + 1)
        per_value = np.zeros_like(count)
        np.add.at(per_value, vals, cells[rings.region[ext]])
        assert np.array_equal(per_value, count)
        # consecutive vertices differ in exactly one coordinate (corners only, axis-parallel edges)
        r, c = rings.vrow.astype(np.int64), rings.vcol.astype(np.int64)
        nxt = np.arange(len(r)) + 1
        nxt[rings.offset[1:] - 1] = rings.offset[:-1]
        assert np.all((r != r[nxt]) ^ (c != c[nxt]))
        assert rings.nedges == int(np.abs(r - r[nxt]).sum() + np.abs(c - c[nxt]).sum())
        del per_region


def test_vectorize_labels_file(tmp_path):
    """vector.py:42-87 end to end: GeoTIFF -> device decode -> rings -> GeoJSON features in world coordinates"""
    from malstroem_b200 import io as mio
    lab = np.load(GOLDEN)["labelled"]
    path = str(tmp_path / "labelled.tif")
    mio.RasterWriter(path, TRANSFORM, "EPSG:25832").write(lab)
    feats = list(vector.vectorize_labels_file(path, "bspot_id"))
    assert len(feats) == 113
    ref = {(p["value"], p["region"]): p for p in P.polygonize(lab)}
    by_first = sorted(ref.values(), key=lambda p: p["region"])
    for f, p in zip(feats, by_first):
        assert f["properties"] == {"bspot_id": p["value"]}
        assert len(f["geometry"]["coordinates"]) == len(p["rings"])
        xs = [pt[0] for pt in f["geometry"]["coordinates"][0]]
        ys = [pt[1] for pt in f["geometry"]["coordinates"][0]]
        cols = [c for _, c in p["rings"][0]]
        rows = [r for r, _ in p["rings"][0]]
        assert min(xs) == TRANSFORM[0] + min(cols) * TRANSFORM[1] and max(xs) == TRANSFORM[0] + max(cols) * TRANSFORM[1]
        assert max(ys) == TRANSFORM[3] + min(rows) * TRANSFORM[5] and min(ys) == TRANSFORM[3] + max(rows) * TRANSFORM[5]
