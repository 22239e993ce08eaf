"""the raster half of malstroem/vector.py on the device.

`vectorize_labels_file` (vector.py:42-87) runs `gdal.Polygonize(band, band.GetMaskBand(), layer, 0, ['8CONNECTED=8'])`
on a label raster and yields one GeoJSON feature per polygon.  Here the file is decoded on the device (io.py), the
rings are built by csrc/polygon.cu (ms_polygonize*: boundary edges, ring leaders and positions by pointer doubling,
corner compaction, regions by union-find) and this module groups rings into features and applies the geotransform.

Same features as GDAL's: one polygon per 8-connected region of equal value (value 0 included: the reference's fixture
gives 113 = 104 labels + 9 background regions, tests/test_vector.py:18-20), holes as inner rings, vertices at pixel
corners in world coordinates, only where the outline turns.  NOT pinned against GDAL (absent here): the order of the
features (here: by the region's first cell in raster order; GDAL: by the row that completes the polygon), the start
vertex and the winding of each ring.  What the tests pin instead: the reference's own feature count, exact equality of
the rasterised features with the raster, and ring-for-ring equality with an independent CPU walk (oracle/polygonize.py).
"""
import numpy as np

from . import _lib


def transform_cell_to_world(cell, geotransform):
    """vector.py:21-39 — world coordinates of the centre of cell (row, col); `gdal.ApplyGeoTransform`'s arithmetic."""
    row, col = cell[:2]
    px, ln = col + 0.5, row + 0.5
    x = geotransform[0] + px * geotransform[1] + ln * geotransform[2]
    y = geotransform[3] + px * geotransform[4] + ln * geotransform[5]
    return (x, y)


class Rings(object):
    """Rings of a polygonised raster: ring k owns vertices offset[k]:offset[k+1] of (vrow, vcol) — lattice corners,
    the region on the right-hand side walking the ring on the screen, no repeated closing vertex."""

    def __init__(self, shape, offset, value, cell, region, hole, vrow, vcol, nedges):
        self.shape = shape
        self.offset, self.value, self.cell, self.region, self.hole = offset, value, cell, region, hole
        self.vrow, self.vcol = vrow, vcol
        self.nedges = nedges

    def __len__(self):
        return len(self.value)

    def polygons(self):
        """[(value, [exterior ring index, hole ring index, ...]), ...] ordered by the region's first cell."""
        ext = np.flatnonzero(self.hole == 0)
        holes = np.flatnonzero(self.hole != 0)
        # an exterior ring's leader cell IS its region's first cell, and the rings come sorted by leader cell
        ext_region = self.region[ext]
        if len(ext) > 1 and not np.all(np.diff(ext_region) > 0):
            raise RuntimeError("polygonize: exterior rings are not one per region")
        where = np.searchsorted(ext_region, self.region[holes])
        if len(holes) and (np.any(where >= len(ext)) or np.any(ext_region[np.minimum(where, len(ext) - 1)] != self.region[holes])):
            raise RuntimeError("polygonize: a hole without an exterior ring")
        order = np.argsort(where, kind="stable")
        first = np.searchsorted(where[order], np.arange(len(ext) + 1))
        out = []
        for k in range(len(ext)):
            out.append((int(self.value[ext[k]]), [int(ext[k])] + [int(h) for h in holes[order[first[k]:first[k + 1]]]]))
        return out

    def ring_lattice(self, k):
        a, b = int(self.offset[k]), int(self.offset[k + 1])
        return self.vrow[a:b], self.vcol[a:b]

    def features(self, geotransform=None, id_attribute="bspot_id"):
        """GeoJSON features as OGR's Feature.ExportToJson gives them (vector.py:78-80)."""
        gt = geotransform if geotransform is not None else (0.0, 1.0, 0.0, 0.0, 0.0, 1.0)
        c, r = self.vcol.astype(np.float64), self.vrow.astype(np.float64)
        x = gt[0] + c * gt[1] + r * gt[2]
        y = gt[3] + c * gt[4] + r * gt[5]
        for fid, (value, rings) in enumerate(self.polygons()):
            coords = []
            for k in rings:
                a, b = int(self.offset[k]), int(self.offset[k + 1])
                ring = [[float(x[i]), float(y[i])] for i in range(a, b)]
                ring.append(ring[0])
                coords.append(ring)
            yield {"type": "Feature", "geometry": {"type": "Polygon", "coordinates": coords},
                   "properties": {id_attribute: value}, "id": fid}


def _fetch(shape, counts):
    nr, nv, ne = int(counts[0]), int(counts[1]), int(counts[2])
    offset = np.zeros(nr + 1, dtype=np.int64)
    value = np.zeros(nr, dtype=np.int32)
    cell = np.zeros(nr, dtype=np.int64)
    region = np.zeros(nr, dtype=np.int64)
    hole = np.zeros(nr, dtype=np.uint8)
    vrow = np.zeros(nv, dtype=np.int32)
    vcol = np.zeros(nv, dtype=np.int32)
    _lib.check(_lib.lib().ms_polygonize_fetch(_lib.ptr(offset), _lib.ptr(value), _lib.ptr(cell), _lib.ptr(region),
                                              _lib.ptr(hole), _lib.ptr(vrow), _lib.ptr(vcol)), "polygonize")
    return Rings(shape, offset, value, cell, region, hole, vrow, vcol, ne)


def _nodata_args(nodata):
    if nodata is None:
        return 0, 0
    if float(nodata) != int(nodata) or not (-2 ** 31 <= int(nodata) < 2 ** 31):
        return 0, 0          # no int32 cell can equal it
    return 1, int(nodata)


def polygonize_labels(labels, connect8=True, nodata=None):
    """Rings of a 2-D int32 label raster (numpy).  gdal.Polygonize's region rule; see the module docstring."""
    if not isinstance(labels, np.ndarray) or labels.ndim != 2:
        raise ValueError("polygonize_labels: a 2-D numpy array is required")
    if labels.dtype != np.int32:
        if not np.issubdtype(labels.dtype, np.integer) and labels.dtype != np.bool_:
            raise ValueError("Buffer dtype mismatch, expected an integer raster but got %s" % labels.dtype)
        labels = labels.astype(np.int32)
    labels = np.ascontiguousarray(labels)
    if labels.size == 0:
        raise ValueError("Width or height of processing area is zero")
    has, nd = _nodata_args(nodata)
    counts = np.zeros(3, dtype=np.int64)
    with _lib.lock:
        _lib.check(_lib.lib().ms_polygonize(_lib.ptr(labels), labels.shape[0], labels.shape[1], 1 if connect8 else 0,
                                            has, nd, _lib.ptr(counts)), "polygonize")
        _lib.track(labels)
        return _fetch(labels.shape, counts)


def polygonize_labels_device(labels, connect8=True, nodata=None):
    """The same for a 2-D int32 CUDA tensor (e.g. RasterReader.read_device or RasterPipeline's labels)."""
    import torch
    if labels.dim() != 2 or labels.dtype != torch.int32 or not labels.is_cuda:
        raise ValueError("polygonize_labels_device: a 2-D int32 CUDA tensor is required")
    labels = labels.contiguous()
    has, nd = _nodata_args(nodata)
    counts = np.zeros(3, dtype=np.int64)
    stream = torch.cuda.current_stream(labels.device).cuda_stream
    with _lib.lock:
        _lib.check(_lib.lib().ms_polygonize_dev(labels.data_ptr(), labels.shape[0], labels.shape[1],
                                                1 if connect8 else 0, has, nd, _lib.ptr(counts), stream), "polygonize")
        return _fetch(tuple(labels.shape), counts)


def vectorize_labels_file(labeled_file, id_attribute="bspot_id"):
    """vector.py:42-87 — yields one GeoJSON feature per polygon of the label raster in `labeled_file` (GeoTIFF)."""
    from . import io
    reader = io.RasterReader(labeled_file)
    labels = reader.read_device()
    import torch
    if labels.dtype != torch.int32:
        if labels.dtype.is_floating_point:
            raise ValueError("vectorize_labels_file: an integer raster is required, got %s" % labels.dtype)
        labels = labels.to(torch.int32)
    rings = polygonize_labels_device(labels, connect8=True, nodata=reader.nodata)
    del labels
    for feature in rings.features(reader.transform, id_attribute):
        yield feature
