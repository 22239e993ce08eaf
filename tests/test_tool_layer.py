"""BASELINE config 1: the reference's own tool layer (`malstroem complete`: DemTool -> BluespotTool -> StreamTool ->
RainTool, scripts/complete.py:57-127) running UNCHANGED on top of malstroem_b200.speedups.enable(), compared with the
same chain on the reference's own compiled Cython path, and with the known answers of the reference's
tests/test_commandline.py:10-47 (486 bluespots / 544 events with the filter, 523 / 587 without).  File I/O (GDAL) is
replaced by in-memory readers / writers (tests/toolchain.py)."""
import numpy as np
import pytest

import toolchain

needs_ref = pytest.mark.skipif(not toolchain.available(), reason="baseline/_ref not installed (baseline/install_ref.py)")
FILTER = 'area > 20.5 and maxdepth > 0.5 or volume > 2.5'      # tests/test_commandline.py:15


def _props(features, drop=()):
    return [{k: v for k, v in f["properties"].items() if k not in drop} for f in features]


def _check_known_answers(out, nlabels, nevents):
    assert int(np.max(out["bluespots"])) == nlabels
    assert len(out["events"]) == nevents
    assert len(out["pourpoints"]) == nlabels + 1             # label 0 gets a pour point too (bluespots.py:74)


@needs_ref
@pytest.mark.parametrize("filt,nlabels,nevents", [(FILTER, 486, 544), (None, 523, 587)])
def test_reference_chain_known_answers_cpu(dtm188, filt, nlabels, nevents):
    """Scaffolding check (no GPU): the in-memory chain on the reference's own Cython path reproduces the reference's
    command-line known answers and golden rasters."""
    alg, *_ = toolchain.import_reference()
    alg.speedups.enable()
    assert alg.speedups.enabled
    out = toolchain.run_complete(dtm188["dtm"], [10, 100], filt)
    _check_known_answers(out, nlabels, nevents)
    assert np.array_equal(out["filled"], dtm188["filled"]) and np.array_equal(out["flowdir"], dtm188["flowdir_noflats"])
    assert np.array_equal(out["depths"], dtm188["depths"])


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("filt,accum,nlabels,nevents", [(FILTER, False, 486, 544), (None, False, 523, 587),
                                                        ('maxdepth > 0.05', True, None, None)])
def test_complete_on_b200_path_equals_reference(dtm188, filt, accum, nlabels, nevents):
    from malstroem_b200 import speedups
    from malstroem_b200.algorithms import fill
    alg, *_ = toolchain.import_reference()
    alg.speedups.enable()                                     # the reference's own native path
    ref = toolchain.run_complete(dtm188["dtm"], [10, 30, 100], filt, accum=accum)
    speedups.enable()
    try:
        assert alg.fill.fill_terrain is fill.fill_terrain and alg.speedups.enabled
        ours = toolchain.run_complete(dtm188["dtm"], [10, 30, 100], filt, accum=accum)
    finally:
        speedups.disable()
    if nlabels is not None:
        _check_known_answers(ours, nlabels, nevents)
    for k in ("filled", "flowdir", "depths", "bluespots", "watersheds") + (("accum",) if accum else ()):
        assert ours[k].dtype == ref[k].dtype and np.array_equal(ours[k], ref[k]), k
    assert np.array_equal(ours["filled"], dtm188["filled"]) and np.array_equal(ours["flowdir"], dtm188["flowdir_noflats"])
    # pour points: cells and counts exact, volumes within 1e-6 relative (north_star)
    vol = ("bspot_vol", "bspot_fumm")
    assert _props(ours["pourpoints"], vol) == _props(ref["pourpoints"], vol)
    for key in vol:
        a = np.array([f["properties"][key] for f in ours["pourpoints"]], dtype=float)
        b = np.array([f["properties"][key] for f in ref["pourpoints"]], dtype=float)
        np.testing.assert_allclose(a, b, rtol=1e-6, atol=0, equal_nan=True)
    # nodes (junction nodes included) and stream geometries: identical
    assert _props(ours["nodes"], ("bspot_vol",)) == _props(ref["nodes"], ("bspot_vol",))
    assert [f["geometry"] for f in ours["nodes"]] == [f["geometry"] for f in ref["nodes"]]
    assert ours["streams"] == ref["streams"]
    # rain events per node id: volumes within 1e-6 relative
    ev_o = {f["properties"]["nodeid"]: f["properties"] for f in ours["events"]}
    ev_r = {f["properties"]["nodeid"]: f["properties"] for f in ref["events"]}
    assert ev_o.keys() == ev_r.keys()
    for nid, pr in ev_r.items():
        for mm in (10, 30, 100):
            for q in ("rainv", "spillv", "v", "pctv"):
                key = "%s_%g" % (q, mm)
                a, b = ev_o[nid][key], pr[key]
                assert (a is None) == (b is None), (nid, key)
                if b is not None:
                    assert abs(a - b) <= 1e-6 * abs(b) + 1e-9, (nid, key, a, b)


@needs_ref
@pytest.mark.gpu
def test_bluespot_tool_vector_outputs(dtm188, tmp_path):
    """§8(f4): the unchanged BluespotTool with its vector outputs on (bluespots.py:177-193): the polygons come from
    the device path, one per 8-connected region of the rasters it wrote, and rasterise back to exactly those rasters"""
    from malstroem_b200 import speedups
    from oracle import polygonize as P
    toolchain.import_reference()
    speedups.enable()
    try:
        out = toolchain.run_bluespots_with_vectors(dtm188["depths"], dtm188["flowdir_noflats"], dtm188["dtm"], str(tmp_path))
    finally:
        speedups.disable()
    gt = toolchain.DTM188_TRANSFORM
    for raster, feats, key in ((out["bluespots"], out["bluespots_vector"], "bspot_id"),
                               (out["watersheds"], out["watersheds_vector"], "bspot_id")):
        ref = P.polygonize(raster, nodata=0)          # the label writers carry nodata 0: no polygon for the background
        assert len(feats) == len(ref) and len(feats) >= int(raster.max())
        polys = []
        for f in feats:
            rings = []
            for ring in f["geometry"]["coordinates"]:
                assert ring[0] == ring[-1]
                rings.append([(int(round((y - gt[3]) / gt[5])), int(round((x - gt[0]) / gt[1]))) for x, y in ring[:-1]])
            polys.append({"value": f["properties"][key], "region": 0, "rings": rings})
        back, twice, _ = P.rasterize(polys, raster.shape, 0)
        assert twice == 0 and np.array_equal(back, raster)
    assert int(out["bluespots"].max()) == 523         # no filter: tests/test_commandline.py:43
