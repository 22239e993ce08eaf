"""§8(f4) — the polygonisation oracle pinned to the reference's known answer, and the host half of
malstroem_b200/vector.py (no GPU needed)."""
import os

import numpy as np
import pytest
from scipy import ndimage

from oracle import polygonize as P
from poly_cases import cases

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "dtm188.npz")


def _region_count(a, connect8=True, nodata=None):
    s = np.ones((3, 3)) if connect8 else None
    return sum(ndimage.label(a == v, structure=s)[1] for v in np.unique(a) if nodata is None or v != nodata)


def test_reference_feature_count():
    """/root/reference/tests/test_vector.py:18-20: 113 features for tests/data/labelled.tif (copy in dtm188.npz)"""
    lab = np.load(GOLDEN)["labelled"]
    polys = P.polygonize(lab)
    assert len(polys) == 113
    assert sorted(set(p["value"] for p in polys)) == list(range(0, 105))
    ras, twice, _ = P.rasterize(polys, lab.shape, -1)
    assert twice == 0 and np.array_equal(ras, lab)


def test_watersheds_fixture():
    ws = np.load(GOLDEN)["wsheds"]
    polys = P.polygonize(ws)
    assert len(polys) == _region_count(ws)
    ras, twice, _ = P.rasterize(polys, ws.shape, -1)
    assert twice == 0 and np.array_equal(ras, ws)


@pytest.mark.parametrize("name", sorted(cases()))
@pytest.mark.parametrize("connect8", [True, False])
def test_oracle_round_trip(name, connect8):
    a = cases()[name]
    polys = P.polygonize(a, connect8=connect8)
    assert len(polys) == _region_count(a, connect8)
    ras, twice, _ = P.rasterize(polys, a.shape, -(2 ** 40))
    assert twice == 0 and np.array_equal(ras, a)
    for p in polys:
        assert P.area2(p["rings"][0]) > 0 and all(P.area2(h) < 0 for h in p["rings"][1:])
        cells = sum(P.area2(g) for g in p["rings"]) // 2
        assert cells > 0


def test_oracle_nodata():
    a = cases()["random2"]
    polys = P.polygonize(a, nodata=0)
    assert all(p["value"] != 0 for p in polys) and len(polys) == _region_count(a, True, 0)
    ras, twice, _ = P.rasterize(polys, a.shape, 0)
    assert twice == 0 and np.array_equal(ras, a)


def test_transform_cell_to_world():
    """/root/reference/tests/test_vector.py:6-15"""
    from malstroem_b200 import vector
    gt = (720000.0, 0.4, 0.0, 6193000.0, 0.0, -0.4)
    assert vector.transform_cell_to_world((0, 0), gt) == (720000.2, 6192999.8)
    assert vector.transform_cell_to_world((1, 10), gt) == (720004.2, 6192999.4)


def test_rings_grouping_and_features():
    """Rings.polygons / features on hand-made arrays: holes follow their exterior ring, rings are closed, the
    geotransform is applied to lattice corners"""
    from malstroem_b200.vector import Rings
    # a 3x3 raster of 1 with a 0 in the middle: exterior of 1 (leader cell 0), hole of 1 (leader cell 1, region 0),
    # exterior of the 0 (cell 4)
    offset = np.array([0, 4, 8, 12], dtype=np.int64)
    vrow = np.array([0, 0, 3, 3, 2, 2, 1, 1, 1, 1, 2, 2], dtype=np.int32)
    vcol = np.array([0, 3, 3, 0, 2, 1, 1, 2, 1, 2, 2, 1], dtype=np.int32)
    r = Rings((3, 3), offset, np.array([1, 1, 0], dtype=np.int32), np.array([0, 1, 4]), np.array([0, 0, 4]),
              np.array([0, 1, 0], dtype=np.uint8), vrow, vcol, 16)
    assert r.polygons() == [(1, [0, 1]), (0, [2])]
    feats = list(r.features((100.0, 2.0, 0.0, 50.0, 0.0, -2.0), "wshed_id"))
    assert [f["properties"] for f in feats] == [{"wshed_id": 1}, {"wshed_id": 0}]
    assert feats[0]["geometry"]["coordinates"][0] == [[100.0, 50.0], [106.0, 50.0], [106.0, 44.0], [100.0, 44.0], [100.0, 50.0]]
    assert len(feats[0]["geometry"]["coordinates"]) == 2 and feats[0]["id"] == 0 and feats[1]["id"] == 1


def test_oracle_round_trip_property():
    """hypothesis: any small label raster, both connectivities, with and without a nodata value"""
    from hypothesis import given, settings, strategies as st
    from hypothesis.extra import numpy as hnp

    @settings(max_examples=120, deadline=None)
    @given(hnp.arrays(np.int32, st.tuples(st.integers(1, 9), st.integers(1, 9)), elements=st.integers(0, 3)),
           st.booleans(), st.sampled_from([None, 0, 2]))
    def check(a, connect8, nodata):
        polys = P.polygonize(a, connect8=connect8, nodata=nodata)
        assert len(polys) == _region_count(a, connect8, nodata)
        fill = -7 if nodata is None else nodata
        ras, twice, _ = P.rasterize(polys, a.shape, fill)
        assert twice == 0 and np.array_equal(ras, a)
        # one exterior ring per polygon, holes negative, and the areas add up to the cells outside the mask
        cells = sum(sum(P.area2(g) for g in p["rings"]) // 2 for p in polys)
        assert cells == int((a != nodata).sum()) if nodata is not None else cells == a.size

    check()
