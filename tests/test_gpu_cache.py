"""Device twins of host rasters across plug-in calls (csrc/cache.cu): the call sequence of
DemTool.process / BluespotTool.process (dem.py:67-91, bluespots.py:158-206) through the numpy mirrors must give the
oracle's results with the cache on and off, reuse what is on the device (hits counted by the library), and never
serve a stale twin for an array that died or was rewritten."""
import gc

import numpy as np
import pytest

from malstroem_b200 import _lib, synth
from malstroem_b200.algorithms import fill, flow, label
from oracle import port

pytestmark = pytest.mark.gpu


def tool_sequence(dem):
    """The twelve calls in the order and on the operands the reference's tools use."""
    filled = fill.fill_terrain(dem)
    depths = filled - dem
    del filled
    short, diag = fill.minimum_safe_short_and_diag(dem)
    fnf = fill.fill_terrain_no_flats(dem, short=short, diag=diag)
    flowdir = flow.terrain_flowdirection(fnf, edges_flow_outward=True)
    del fnf
    accum = flow.accumulated_flow(flowdir)
    raw, nraw = label.connected_components(depths)
    raw_stats = label.label_stats(depths, raw)
    keep = [bool(m > 0.05) for m in raw_stats["max"]]
    comps = label.keep_labels(raw, keep)
    del raw
    lab, n = label.connected_components(comps)
    stats = label.label_stats(depths, lab)
    ws = np.copy(lab)
    flow.watersheds_from_labels(flowdir, ws, unassigned=0)
    wcount = label.label_count(ws)
    ppmax = label.label_max_index(accum, lab, n)
    short2, diag2 = fill.minimum_safe_short_and_diag(dem)
    fnf2 = fill.fill_terrain_no_flats(dem, short2, diag2)
    ppmin = label.label_min_index(fnf2, lab, n)
    return dict(depths=depths, flowdir=flowdir, accum=accum, lab=lab, n=n, stats=stats, ws=ws, wcount=wcount,
                ppmax=ppmax, ppmin=ppmin, fnf=fnf2, nraw=nraw)


def oracle_sequence(dem):
    filled = port.fill_terrain(dem)
    depths = filled - dem
    short, diag = port.minimum_safe_short_and_diag(dem)
    fnf = port.fill_terrain_no_flats(dem, short, diag)
    flowdir = port.terrain_flowdirection(fnf)
    accum = port.accumulated_flow(flowdir, fast=True)
    raw, nraw = port.connected_components(depths)
    raw_stats = port.label_stats(depths, raw, nraw)
    keep = raw_stats["max"] > 0.05
    keep[0] = False
    lab, n = port.connected_components(keep[raw])
    stats = port.label_stats(depths, lab, n)
    ws = lab.copy()
    port.watersheds_from_labels(flowdir, ws, 0)
    return dict(depths=depths, flowdir=flowdir, accum=accum, lab=lab, n=n, stats=stats, ws=ws,
                wcount=port.label_count(ws), ppmax=port.label_max_index(accum, lab, n),
                ppmin=port.label_min_index(fnf, lab, n), fnf=fnf, nraw=nraw)


def same(got, want):
    for k in ("depths", "flowdir", "accum", "lab", "ws", "fnf", "wcount"):
        assert np.array_equal(got[k], want[k]), k
    assert got["n"] == want["n"] and got["nraw"] == want["nraw"]
    for k in ("min", "max", "count"):
        assert np.array_equal(got["stats"][k], want["stats"][k]), k
    np.testing.assert_allclose(got["stats"]["sum"], want["stats"]["sum"], rtol=1e-6, atol=1e-300)
    for key in ("ppmin", "ppmax"):
        for k in ("value", "row", "col"):
            assert np.array_equal(got[key][k], want[key][k]), (key, k)


def test_tool_sequence_with_cache_equals_oracle_and_reuses():
    dem = synth.fractal_dem(384, 512, seed=21)
    want = oracle_sequence(dem)
    _lib.cache_clear()
    before = _lib.cache_stats()
    got = tool_sequence(dem)
    after = _lib.cache_stats()
    same(got, want)
    # the DEM is uploaded once (fill, min/max x2, no-flats x2 find it), the no-flats fill finds the plain fill and its
    # second call the surface itself
    assert after["derived_hits"] - before["derived_hits"] >= 2
    assert after["input_hits"] - before["input_hits"] >= 12
    uploads = after["input_uploads"] - before["input_uploads"]
    assert uploads <= 4, uploads       # dem, depths (numpy result), the watershed copy (+ nothing else)


def test_cache_off_gives_the_same(monkeypatch):
    import ctypes
    import os
    import subprocess
    import sys
    # the switch is read once per process: run the sequence in a child with MS_CACHE=0
    code = ("import numpy as np, sys; sys.path.insert(0, %r); sys.path.insert(0, %r);"
            "from malstroem_b200 import synth, _lib; from test_gpu_cache import tool_sequence, oracle_sequence, same;"
            "dem = synth.fractal_dem(200, 260, seed=5); same(tool_sequence(dem), oracle_sequence(dem));"
            "st = _lib.cache_stats(); assert st['input_hits'] == 0 and st['derived_hits'] == 0 and st['bytes'] == 0, st;"
            "print('ok')") % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                               os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MS_CACHE="0")
    r = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout[-2000:]


def test_no_stale_twin_after_rewrite_or_death():
    dem = synth.fractal_dem(256, 300, seed=8)
    filled = fill.fill_terrain(dem)
    depths = filled - dem
    lab1, n1 = label.connected_components(depths)
    # the same array object, rewritten wholesale: the fingerprint no longer matches
    depths[...] = np.where(depths > 0.02, depths, 0).astype(np.float32)
    lab2, n2 = label.connected_components(depths)
    want, nw = port.connected_components(depths)
    assert n2 == nw and np.array_equal(lab2, want)
    # arrays that die: their addresses come back with other contents
    for seed in range(6):
        d = synth.fractal_dem(256, 300, seed=30 + seed)
        f = fill.fill_terrain(d)
        assert np.array_equal(f, port.fill_terrain(d))
        del d, f
        gc.collect()
    # explicit clear
    _lib.cache_clear()
    assert _lib.cache_stats()["bytes"] == 0
    assert np.array_equal(fill.fill_terrain(dem), filled)


def test_second_device_is_rejected():
    from malstroem_b200 import _lib as L
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    with pytest.raises(ValueError):
        L.check(L.lib().ms_init(1), "ms_init")
