"""Test scaffolding for driving the REFERENCE's own tool layer (DemTool, BluespotTool, StreamTool, RainTool — what
`malstroem complete` chains, malstroem/scripts/complete.py:57-127) without GDAL: an `osgeo` stand-in that covers
the one call the tools make outside file I/O (gdal.ApplyGeoTransform, vector.py:39), and in-memory readers /
writers with the interface of malstroem/io.py (as /root/reference/tests/test_raster_bluespot.py:9-15 does).
The reference package comes from baseline/_ref (installed by baseline/install_ref.py, unmodified apart from the
dtype-spelling patch its Cython build needs)."""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_INSTALL = os.path.join(ROOT, "baseline", "_ref")
# geo-transform of the reference's tests/data/dtm.tif (GeoTIFF tags 33550 / 33922)
DTM188_TRANSFORM = (720000.0, 16.0, 0.0, 6193000.0, 0.0, -15.957446808510639)


def available():
    return os.path.isdir(os.path.join(REF_INSTALL, "malstroem"))


def _install_osgeo_stub():
    if "osgeo" in sys.modules:
        return
    osgeo = types.ModuleType("osgeo")
    gdal = types.ModuleType("osgeo.gdal")
    ogr = types.ModuleType("osgeo.ogr")
    osr = types.ModuleType("osgeo.osr")

    def apply_geotransform(gt, x, y):       # the documented 6-parameter affine
        return gt[0] + x * gt[1] + y * gt[2], gt[3] + x * gt[4] + y * gt[5]

    gdal.ApplyGeoTransform = apply_geotransform
    for k, name in enumerate(("wkbPoint", "wkbLineString", "wkbPolygon", "wkbMultiPolygon")):
        setattr(ogr, name, k + 1)
    osgeo.gdal, osgeo.ogr, osgeo.osr = gdal, ogr, osr
    sys.modules.update({"osgeo": osgeo, "osgeo.gdal": gdal, "osgeo.ogr": ogr, "osgeo.osr": osr})


def import_reference():
    """Put baseline/_ref first on sys.path and import the reference's tool modules."""
    _install_osgeo_stub()
    if REF_INSTALL not in sys.path:
        sys.path.insert(0, REF_INSTALL)
    for k in [k for k in sys.modules if k == "malstroem" or k.startswith("malstroem.")]:
        if not getattr(sys.modules[k], "__file__", "") or not str(sys.modules[k].__file__).startswith(REF_INSTALL):
            del sys.modules[k]
    import malstroem.algorithms as alg
    from malstroem import bluespots, dem, rain, streams
    return alg, dem, bluespots, streams, rain


class MemRaster(object):
    """RasterReader and RasterWriter of malstroem/io.py in one: write() stores, read() returns."""

    def __init__(self, data=None, transform=DTM188_TRANSFORM):
        self.data = data
        self.transform = list(transform)
        self.crs = None
        self.filepath = None

    def read(self):
        return self.data

    def write(self, data):
        self.data = data


class MemVector(object):
    """VectorReader / VectorWriter of malstroem/io.py: geojson features in, the same features out."""

    def __init__(self):
        self.features = []

    def write_geojson_features(self, features):
        if isinstance(features, dict):
            features = features["features"]
        self.features = list(features)

    def read_geojson_features(self):
        return self.features


def parse_filter(text):
    # malstroem/scripts/_utils.py:29-40 (the CLI's -filter option)
    if not text:
        return lambda stats: True
    expr = text.replace('area', 'stats["area"]').replace('maxdepth', 'stats["max"]').replace('volume', 'stats["volume"]')
    return eval('lambda stats: {}'.format(expr))


class FileRaster(MemRaster):
    """A writer that also leaves a GeoTIFF behind (BluespotTool vectorises from `filepath`, bluespots.py:179,192)."""

    def __init__(self, filepath, transform=DTM188_TRANSFORM):
        MemRaster.__init__(self, None, transform)
        self.filepath = filepath

    def write(self, data):
        from malstroem_b200 import io as mio
        self.data = data
        # nodata 0 as scripts/complete.py:83-85 and scripts/bluespot.py:44,79 create the label writers
        mio.RasterWriter(self.filepath, tuple(self.transform), "EPSG:25832", 0).write(data)


def run_bluespots_with_vectors(depths, flowdir, dem_array, tmpdir, filter_text=None, transform=DTM188_TRANSFORM):
    """BluespotTool with the `-vector` outputs of `malstroem bluespots` (scripts/bluespot.py) switched on."""
    _, _, bluespots, _, _ = import_reference()
    labeled = FileRaster(os.path.join(tmpdir, "bluespots.tif"), transform)
    wsheds = FileRaster(os.path.join(tmpdir, "wsheds.tif"), transform)
    labeled_vec, wsheds_vec, pourpoints = MemVector(), MemVector(), MemVector()
    bluespots.BluespotTool(input_depths=MemRaster(depths, transform), input_flowdir=MemRaster(flowdir, transform),
                           input_bluespot_filter_function=parse_filter(filter_text), input_accum=None,
                           input_dem=MemRaster(dem_array, transform), output_labeled_raster=labeled,
                           output_labeled_vector=labeled_vec, output_pourpoints=pourpoints,
                           output_watersheds_raster=wsheds, output_watersheds_vector=wsheds_vec).process()
    return dict(bluespots=labeled.data, watersheds=wsheds.data, bluespots_vector=labeled_vec.features,
                watersheds_vector=wsheds_vec.features)


def run_complete(dem_array, rain_mm, filter_text=None, accum=False, transform=DTM188_TRANSFORM, streams_geometry=True):
    """The tool sequence of `malstroem complete` (scripts/complete.py:57-127) on in-memory rasters.  Returns dict of
    every raster and feature list the command would write."""
    _, dem, bluespots, streams, rain = import_reference()
    dem_reader = MemRaster(dem_array, transform)
    filled, flowdir, depths = MemRaster(None, transform), MemRaster(None, transform), MemRaster(None, transform)
    accum_w = MemRaster(None, transform) if accum else None
    dem.DemTool(dem_reader, filled, flowdir, depths, accum_w).process()
    pourpoints, wsheds, labeled = MemVector(), MemRaster(None, transform), MemRaster(None, transform)
    bluespots.BluespotTool(input_depths=depths, input_flowdir=flowdir,
                           input_bluespot_filter_function=parse_filter(filter_text), input_accum=accum_w,
                           input_dem=dem_reader, output_labeled_raster=labeled, output_labeled_vector=None,
                           output_pourpoints=pourpoints, output_watersheds_raster=wsheds,
                           output_watersheds_vector=None).process()
    nodes, stream_lines = MemVector(), (MemVector() if streams_geometry else None)
    streams.StreamTool(pourpoints, labeled, flowdir, nodes, stream_lines).process()
    events = MemVector()
    rain.RainTool(nodes, events, rain_mm).process()
    return dict(filled=filled.data, flowdir=flowdir.data, depths=depths.data, accum=accum_w.data if accum else None,
                bluespots=labeled.data, watersheds=wsheds.data, pourpoints=pourpoints.features, nodes=nodes.features,
                streams=stream_lines.features if streams_geometry else None, events=events.features)
