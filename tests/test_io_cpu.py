"""Host side of the raster codec (malstroem_b200/io.py): the TIFF container.  The tag directory
this package writes parses back to what went in (classic and BigTIFF), and the tags of the reference's own
GDAL-written fixtures are read correctly (no GPU: nothing is decoded here)."""
import os
import struct

import numpy as np
import pytest

from malstroem_b200 import io as mio

TIF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tif")
TRANSFORM = (720000.0, 16.0, 0.0, 6193000.0, 0.0, -15.957446808510639)


@pytest.mark.parametrize("dtype,predictor", [(np.float32, 2), (np.float64, 1), (np.int32, 2), (np.uint8, 2)])
def test_container_round_trip(tmp_path, dtype, predictor):
    rows, cols = 300, 517
    sizes = [100 + 7 * k for k in range(2 * 3)]
    head, tail = mio.build_tiff(rows, cols, dtype, predictor, sizes, TRANSFORM, "EPSG:25832 (test)", nodata=-999)
    p = str(tmp_path / "x.tif")
    with open(p, "wb") as f:
        f.write(head + b"".join(bytes([k]) * s for k, s in enumerate(sizes)) + tail)
    t = mio.TiffInfo(p)
    assert (t.rows, t.cols, t.dtype, t.predictor, t.compression) == (rows, cols, np.dtype(dtype), predictor, 8)
    assert (t.block_w, t.block_h) == (256, 256) and t.counts == sizes
    raw = open(p, "rb").read()
    for k, (o, c) in enumerate(zip(t.offsets, t.counts)):
        assert raw[o:o + c] == bytes([k]) * c
    assert t.nodata == -999.0 and t.crs == "EPSG:25832 (test)"
    np.testing.assert_allclose(t.transform, TRANSFORM)


def test_bigtiff_directory(tmp_path):
    # sizes that add up to more than 4 GB switch the container to BigTIFF; only the directory is written here
    sizes = [1 << 30] * 5
    head, tail = mio.build_tiff(70000, 70000, np.float32, 2, sizes, None, None)
    assert head[:4] == b"II" + struct.pack("<H", 43)
    ifd_off = struct.unpack("<Q", head[8:16])[0]
    assert ifd_off == 16 + sum(sizes)
    n = struct.unpack("<Q", tail[:8])[0]
    tags = [struct.unpack("<H", tail[8 + 20 * k:10 + 20 * k])[0] for k in range(n)]
    assert tags == sorted(tags) and {256, 257, 322, 323, 324, 325}.issubset(tags)


@pytest.mark.parametrize("name,dtype,predictor,tiled", [("labelled", np.int32, 2, True), ("flowdir_noflats", np.uint8, 2, True),
                                                        ("filled_no_flats", np.float64, 1, True), ("dtm", np.float32, 2, False)])
def test_reference_fixture_tags(name, dtype, predictor, tiled):
    t = mio.TiffInfo(os.path.join(TIF, name + ".tif"))
    assert (t.rows, t.cols) == (188, 250) and t.dtype == np.dtype(dtype)
    assert t.compression == 8 and t.predictor == predictor
    assert (t.block_w, t.block_h) == ((256, 256) if tiled else (250, 8))
    assert len(t.offsets) == (1 if tiled else 24)
    np.testing.assert_allclose(t.transform, TRANSFORM, rtol=1e-12)


def test_reader_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mio.RasterReader(os.path.join(TIF, "labelled.tif")).read()
