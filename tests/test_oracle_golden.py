"""Pin the CPU oracle (oracle/ms_oracle.c) to the reference: its golden rasters, and outputs of the
reference itself (pure-Python path and compiled Cython modules) stored by tests/golden/make_golden.py.
Mirrors /root/reference/tests/test_raster_{fill,flowdir,label}.py."""
import numpy as np

from oracle import port


def eq(a, b):
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b)


def test_fill_golden(dtm188):                       # tests/test_raster_fill.py:7-13,26-33
    assert eq(port.fill_terrain(dtm188["dtm"]), dtm188["filled"])
    assert eq(dtm188["filled"] - dtm188["dtm"], dtm188["depths"])


def test_fill_no_flats_golden(dtm188):              # tests/test_raster_fill.py:16-23,36-43
    short, diag = port.minimum_safe_short_and_diag(dtm188["dtm"])
    assert short == dtm188["short"] and diag == dtm188["diag"]
    assert diag / short - 2 ** 0.5 < 0.0001          # :70-72
    assert eq(port.fill_terrain_no_flats(dtm188["dtm"], short, diag), dtm188["filled_no_flats"])


def test_negative_dem_values():                     # tests/test_raster_fill.py:75-82
    dtm = np.full((10, 10), -9999, np.float32)
    dtm[4:6, 4:6] = 0
    short, diag = port.minimum_safe_short_and_diag(dtm)
    assert port.fill_terrain_no_flats(dtm, short, diag)[1, 1] != -9999


def test_flowdir_golden(dtm188):                    # tests/test_raster_flowdir.py:12-26
    assert eq(port.terrain_flowdirection(dtm188["filled_no_flats"]), dtm188["flowdir_noflats"])


def test_accum_golden(dtm188):                      # tests/test_raster_flowdir.py:49-70
    for fast in (False, True):
        acc = port.accumulated_flow(dtm188["flowdir_noflats"], fast=fast)
        assert acc.min() >= 1 and acc.max() == 11158 and acc.sum() == 3578615
        assert eq(acc, dtm188["accum"])


def test_watersheds_golden(dtm188):                 # tests/test_raster_flowdir.py:73-140
    for dt in (np.int32, np.int64, np.uint32):
        ws = dtm188["labelled"].astype(dt)
        port.watersheds_from_labels(dtm188["flowdir_noflats"], ws, 0)
        assert ws.dtype == dt and np.array_equal(ws, dtm188["wsheds"]) and ws.sum() == 2337891


def test_connected_components_golden(dtm188):       # tests/test_raster_label.py:8-16
    lab, n = port.connected_components(dtm188["filled_no_flats"] - dtm188["filled"])
    assert lab.dtype == np.int32 and n == 525 and (lab == 0).sum() == 40029 and lab.sum() == 1561377
    assert eq(lab, dtm188["cc_diff_labels"])
    lab, n = port.connected_components(dtm188["depths"])
    assert n == dtm188["raw_nlabels"] and eq(lab, dtm188["raw_labels"])


def test_label_tables_golden(dtm188):               # tests/test_raster_label.py:19-89 + pourpoints.json
    st = port.label_stats(dtm188["depths"], dtm188["labelled"])
    for k in ("min", "max", "sum", "count"):
        assert np.array_equal(st[k], dtm188["lab_stats_" + k])
    mi = port.label_min_index(dtm188["filled_no_flats"], dtm188["labelled"])
    assert np.array_equal(mi["row"], dtm188["pp_cell_row"]) and np.array_equal(mi["col"], dtm188["pp_cell_col"])
    assert np.array_equal(mi["value"], dtm188["lab_minidx_value"])
    ma = port.label_max_index(dtm188["accum"], dtm188["labelled"])
    for k in ("value", "row", "col"):
        assert np.array_equal(ma[k], dtm188["lab_maxidx_" + k])
    assert np.array_equal(port.label_count(dtm188["wsheds"]), dtm188["lab_wshed_count"])
    cell_area = 16.0 * (15.0 / 0.94)                 # pixel 16.0 x 15.957... m
    np.testing.assert_allclose(st["sum"] * cell_area, dtm188["pp_bspot_vol"], rtol=1e-9)
    np.testing.assert_allclose(st["max"], dtm188["pp_bspot_dmax"], rtol=0)


def _check_case(o):
    dem = o["dem"]
    assert eq(port.fill_terrain(dem), o["filled_py"])                      # pure Python = record (F2)
    short, diag = port.minimum_safe_short_and_diag(dem)
    assert short == o["short"] and diag == o["diag"]
    fnf = port.fill_terrain_no_flats(dem, short, diag)
    assert eq(fnf, o["fnf_py"])
    assert eq(port.terrain_flowdirection(fnf, True), o["flowdir"])
    assert eq(port.terrain_flowdirection(fnf, False), o["flowdir_noedge"])
    assert eq(port.accumulated_flow(o["flowdir"]), o["accum"])
    assert eq(port.accumulated_flow(o["flowdir"], fast=True), o["accum"])
    lab, n = port.connected_components(o["depths"])
    assert n == o["nlabels"] and eq(lab, o["labels"])
    ws = lab.copy()
    port.watersheds_from_labels(o["flowdir"], ws, 0)
    assert eq(ws, o["wsheds"])
    assert np.array_equal(port.label_count(ws), o["wshed_count"])
    st = port.label_stats(o["depths"], lab)
    for k in ("min", "max", "sum", "count"):
        assert np.array_equal(st[k], o["stats_" + k])
    mi = port.label_min_index(fnf, lab, n)
    ma = port.label_max_index(o["accum"], lab, n)
    for k in ("value", "row", "col"):
        assert np.array_equal(mi[k], o["minidx_" + k])
        if "maxidx_" + k in o:
            assert np.array_equal(ma[k], o["maxidx_" + k])


def test_small_cases(small_cases):
    cases, flows = small_cases
    assert len(cases) == 40
    for o in cases:
        _check_case(o)
    for f in flows:                                   # NODIR cells, inward/none border directions
        ws = f["labels"].copy()
        port.watersheds_from_labels(f["flowdir"], ws, 0)
        assert eq(ws, f["wsheds"])
        if f["accum"].size:
            assert eq(port.accumulated_flow(f["flowdir"]), f["accum"])


def test_fractal256(fractal256):
    o = dict(fractal256)
    o["filled_py"], o["fnf_py"] = o["filled_cy"], o["fnf_cy"]      # converged there (checked by the generator)
    _check_case(o)


def test_keep_labels():
    lab = np.array([[0, 1, 2], [3, 3, 0]], np.int32)
    keep = [True, False, True, True]
    out = port.keep_labels(lab, keep)
    assert keep[0] is False and out.dtype == bool
    assert np.array_equal(out, np.array([[0, 0, 1], [1, 1, 0]], bool))
