"""the GeoTIFF codec of malstroem/io.py on the device (csrc/tiff.cu + malstroem_b200/io.py).
Parity = the bytes decode to the same arrays:
  * files written by RasterWriter are read back by an independent TIFF reader (cv2 / libtiff, which also checks every
    tile's Adler-32) and by our own reader, for the four dtypes the reference writes, odd shapes and sparse / dense data;
  * the device inflate reads what real zlib writes (stored, fixed and dynamic Huffman blocks: levels 0, 1, 6, 9),
    tiles and strips, predictor 1 and 2, and the reference's own GDAL-written fixtures;
  * RasterReader.read's nodata substitution (io.py:61-72) including its truthiness quirk."""
import os
import zlib

import cv2
import numpy as np
import pytest

from malstroem_b200 import io as mio, synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
TIF = os.path.join(HERE, "golden", "tif")
TRANSFORM = (720000.0, 16.0, 0.0, 6193000.0, 0.0, -15.957446808510639)


def rasters(rows, cols, seed):
    dem = synth.fractal_dem(rows, cols, seed=seed)
    rng = np.random.default_rng(seed)
    lab = np.zeros((rows, cols), np.int32)
    for k in range(12):
        r, c = int(rng.integers(0, max(rows - 8, 1))), int(rng.integers(0, max(cols - 8, 1)))
        lab[r:r + int(rng.integers(3, 60)), c:c + int(rng.integers(3, 90))] = 100000 + 7 * k
    fd = rng.integers(0, 9, (rows, cols)).astype(np.uint8)
    fd[rows // 3:rows // 2] = 4
    f64 = dem.astype(np.float64) + np.arange(cols) * 2.0 ** -37
    return {"f32": dem, "i32": lab, "u8": fd, "f64": f64}


@pytest.mark.parametrize("shape", [(256, 256), (300, 517), (1000, 70), (31, 1025), (1, 5)])
def test_writer_files_decode_with_libtiff_and_with_our_reader(tmp_path, shape):
    for name, arr in rasters(shape[0], shape[1], 3).items():
        p = str(tmp_path / (name + ".tif"))
        w = mio.RasterWriter(p, TRANSFORM, "EPSG:25832", nodata=-999)
        w.write(arr)
        got = cv2.imread(p, cv2.IMREAD_UNCHANGED)
        assert got is not None and got.dtype == arr.dtype and np.array_equal(got.reshape(arr.shape), arr), name
        r = mio.RasterReader(p)
        back = r.read()
        assert back.dtype == arr.dtype and np.array_equal(back, arr), name
        np.testing.assert_allclose(r.transform, TRANSFORM)
        assert r.crs == "EPSG:25832" and r.nodata == -999.0
        assert w.options["tiled"] == "yes" and w.options["compress"] == "deflate"
        assert w.options.get("predictor") == (None if name == "f64" else 2)
    # the label raster is mostly flat: run-length matches must show
    assert w is not None


def test_sparse_rasters_compress():
    import tempfile
    lab = np.zeros((2048, 2048), np.int32)
    lab[100:400, 200:900] = 77
    with tempfile.TemporaryDirectory() as d:
        w = mio.RasterWriter(os.path.join(d, "l.tif"), None, None)
        w.write(lab)
        assert w.stats["file_bytes"] < w.stats["raw_bytes"] / 50
        assert np.array_equal(cv2.imread(os.path.join(d, "l.tif"), cv2.IMREAD_UNCHANGED), lab)


def reference_tiff(path, arr, predictor, level, strips=0):
    """A TIFF whose blocks come from REAL zlib (numpy predictor, zlib.compress): tiles of 256 x 256, or strips."""
    rows, cols = arr.shape
    es = arr.dtype.itemsize
    uint = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[es]
    blocks = []
    if strips:
        for r0 in range(0, rows, strips):
            b = arr[r0:r0 + strips].view(uint).copy()
            if predictor == 2:
                b[:, 1:] = b[:, 1:] - b[:, :-1]
            blocks.append(zlib.compress(b.tobytes(), level))
    else:
        for r0 in range(0, rows, 256):
            for c0 in range(0, cols, 256):
                t = np.zeros((256, 256), uint)
                part = arr[r0:r0 + 256, c0:c0 + 256].view(uint)
                t[:part.shape[0], :part.shape[1]] = part
                if predictor == 2:
                    t[:, 1:] = t[:, 1:] - t[:, :-1]
                blocks.append(zlib.compress(t.tobytes(), level))
    head, tail = mio.build_tiff(rows, cols, arr.dtype, predictor, [len(b) for b in blocks], TRANSFORM, None, None)
    data = head + b"".join(blocks) + tail
    if strips:
        # the same directory with strip tags instead of tile tags (273 / 278 / 279 for 324 / 323 / 325)
        import struct
        ifd_off = struct.unpack("<I", data[4:8])[0]
        n = struct.unpack("<H", data[ifd_off:ifd_off + 2])[0]
        out = bytearray(data)
        for k in range(n):
            o = ifd_off + 2 + 12 * k
            tag = struct.unpack("<H", out[o:o + 2])[0]
            if tag == 324:
                out[o:o + 2] = struct.pack("<H", 273)
            elif tag == 325:
                out[o:o + 2] = struct.pack("<H", 279)
            elif tag == 323:
                out[o:o + 2] = struct.pack("<H", 278)
                out[o + 8:o + 12] = struct.pack("<HH", strips, 0)
            elif tag == 322:
                out[o:o + 2] = struct.pack("<H", 269)      # DocumentName slot: harmless, keeps the directory sorted enough
                out[o + 2:o + 4] = struct.pack("<H", 3)
        data = bytes(out)
    with open(path, "wb") as f:
        f.write(data)


@pytest.mark.parametrize("level", [0, 1, 6, 9])
@pytest.mark.parametrize("strips", [0, 8])
def test_reader_inflates_real_zlib_streams(tmp_path, level, strips):
    for name, arr in rasters(300, 517, 4).items():
        for predictor in (1, 2):
            p = str(tmp_path / ("%s_%d.tif" % (name, predictor)))
            reference_tiff(p, arr, predictor, level, strips)
            back = mio.RasterReader(p).read()
            assert back.dtype == arr.dtype and np.array_equal(back, arr), (name, predictor, level, strips)


def test_reader_decodes_the_references_own_files(dtm188):
    for name, key in (("labelled", "labelled"), ("flowdir_noflats", "flowdir_noflats"), ("dtm", "dtm"),
                      ("filled_no_flats", "filled_no_flats")):
        r = mio.RasterReader(os.path.join(TIF, name + ".tif"))
        got = r.read()
        assert got.dtype == dtm188[key].dtype and np.array_equal(got, dtm188[key]), name
        np.testing.assert_allclose(r.transform, TRANSFORM, rtol=1e-12)


def test_nodata_substitution(tmp_path):
    arr = synth.fractal_dem(260, 300, seed=2)
    arr[5:9, 7:40] = -9999.0
    p = str(tmp_path / "n.tif")
    mio.RasterWriter(p, TRANSFORM, None, nodata=-9999).write(arr)
    want = arr.copy()
    want[np.isclose(want, -9999.0)] = -999
    assert np.array_equal(mio.RasterReader(p, nodatasubst=-999).read(), want)
    assert np.array_equal(mio.RasterReader(p).read(), arr)                     # no substitute given
    # io.py:69 tests the nodata VALUE for truth: a nodata of 0 is never substituted
    z = arr.copy()
    z[z < 20] = 0
    mio.RasterWriter(p, TRANSFORM, None, nodata=0).write(z)
    assert np.array_equal(mio.RasterReader(p, nodatasubst=-999).read(), z)
    # NaN nodata: np.isnan
    q = arr.copy()
    q[100:120, 50:60] = np.nan
    mio.RasterWriter(p, TRANSFORM, None, nodata=float("nan")).write(q)
    got = mio.RasterReader(p, nodatasubst=-999).read()
    want = np.where(np.isnan(q), np.float32(-999), q)
    assert np.array_equal(got, want)


def test_malformed_stream_raises(tmp_path):
    arr = synth.fractal_dem(256, 256, seed=2)
    p = str(tmp_path / "bad.tif")
    reference_tiff(p, arr, 2, 6)
    raw = bytearray(open(p, "rb").read())
    raw[200:260] = b"\xff" * 60
    open(p, "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        mio.RasterReader(p).read()


def test_writer_from_device_tensor_and_big_raster(tmp_path):
    import torch
    from malstroem_b200.pipeline import synth_fractal
    dem = synth_fractal(4096, 4096, seed=1)
    p = str(tmp_path / "big.tif")
    w = mio.RasterWriter(p, TRANSFORM, None)
    w.write(dem)
    assert torch.equal(mio.RasterReader(p).read_device(), dem)
    assert np.array_equal(cv2.imread(p, cv2.IMREAD_UNCHANGED), dem.cpu().numpy())
