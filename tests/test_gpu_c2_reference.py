"""BASELINE configs[1] (synthetic fractal DEM 8192 x 8192, seed 1) — and 1024 / 2048 / 4096 — compared with THE
REFERENCE, not with certificates: tests/golden/c2_hashes.json holds the SHA-256 of every raster and table the
reference's stock functions (baseline/_ref, its own Cython path; accumulation through the pinned C port above 2048^2,
see tests/golden/make_c2_hashes.py) produced on the same DEM in the build container.  Bit-exact for every raster,
count and arg cell; label_stats 'sum' within 1e-6 relative on the stored sample and the total.
The single-GPU device-resident path (RasterPipeline, through the C ABI) and a 4-band run are checked."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from malstroem_b200 import bands, synth
from malstroem_b200.pipeline import RasterPipeline, synth_fractal

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "c2_hashes.json")) as f:
    GOLD = json.load(f)
RASTERS = ("filled", "depths", "fnf", "flowdir", "accum", "labels", "wsheds")
TABLES = ("st_min", "st_max", "st_count", "ws_count", "ppmin_value", "ppmin_row", "ppmin_col", "ppmax_value",
          "ppmax_row", "ppmax_col")


def sha(a):
    a = np.ascontiguousarray(a)
    if a.dtype.kind == "f":
        a = a + 0.0                      # -0.0 -> +0.0 (the one documented deviation)
    return hashlib.sha256(a.tobytes()).hexdigest()


def check(g, rasters, tables, nlabels, short, diag):
    assert nlabels == g["nlabels"]
    assert short == g["short"] and diag == g["diag"]
    for name in rasters:
        assert sha(rasters[name]) == g[name], name
    m = nlabels + 1
    for name in TABLES:
        assert sha(tables[name][:m]) == g[name], name
    s = tables["st_sum"][:m]
    np.testing.assert_allclose(s[:: g["st_sum_stride"]], np.array(g["st_sum_sample"]), rtol=1e-6, atol=1e-300)
    np.testing.assert_allclose(s.sum(), g["st_sum_total"], rtol=1e-6)


@pytest.mark.parametrize("size", [1024, 2048, 4096, 8192])
def test_single_gpu_equals_reference(size):
    g = GOLD[str(size)]
    p = RasterPipeline(size, size, device=0)
    synth_fractal(size, size, seed=g["seed"], out=p.dem)
    if size <= 2048:                     # the device generator is the numpy generator (hash of the DEM itself)
        assert sha(p.dem.cpu().numpy()) == g["dem"] == sha(synth.fractal_dem(size, size, seed=g["seed"]))
    p.run()
    torch.cuda.synchronize()
    rasters = {k: p.out[k].cpu().numpy() for k in RASTERS}
    tables = {k: p.tables[k].cpu().numpy() for k in p.tables}
    check(g, rasters, tables, p.nlabels, p.short, p.diag)


@pytest.mark.parametrize("size,nbands", [(2048, 3), (8192, 4)])
def test_bands_equal_reference(size, nbands):
    g = GOLD[str(size)]
    dem = synth_fractal(size, size, seed=g["seed"])
    pipes = bands.run_threaded(dem, nbands)
    try:
        rasters = {k: torch.cat([p.out[k] for p in pipes]).cpu().numpy() for k in RASTERS}
        for i, p in enumerate(pipes):                    # every rank holds the complete tables
            tables = {k: v.cpu().numpy() for k, v in p.tables.items()}
            check(g, rasters if i == 0 else {}, tables, p.nlabels, p.short, p.diag)
    finally:
        for p in pipes:
            p.close()
