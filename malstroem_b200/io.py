"""Mirror of malstroem/io.py's raster half (RasterReader io.py:21-72, RasterWriter io.py:75-139) without GDAL:
GeoTIFF files — tiled 256 x 256, deflate, predictor 2 for float32 / int32 / uint8, exactly the options the reference's
writer passes to GDAL (io.py:112, :127-136) — are written and read with the codec kernels of csrc/tiff.cu
(SURVEY.md 8(f3)).  The byte-heavy halves (predictor + deflate, inflate + predictor + nodata substitution) run on the
device; this module is the container: header, tag directory, tile offsets (a few hundred bytes per file).

Same constructor arguments, attributes and methods as the reference classes, so `DemTool` / `BluespotTool` accept them
as duck-typed readers / writers.  Differences, all forced by the absence of GDAL's CRS database: `crs` is carried as
the GeoTIFF citation string (written to and read from GTCitationGeoKey) instead of being translated to EPSG geokeys;
readers without GDAL's WKT parser (cv2, tifffile) are unaffected.  The vector half of io.py (OGR) is out of scope.
There is no CPU fallback: read() / write() need the CUDA library.
"""
import ctypes
import struct

import numpy as np

from . import _lib

TILE = 256
# TIFF field types
_BYTE, _ASCII, _SHORT, _LONG, _RATIONAL, _DOUBLE, _LONG8 = 1, 2, 3, 4, 5, 12, 16
_TYPE_SIZE = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 6: 1, 7: 1, 8: 2, 9: 4, 10: 8, 11: 4, 12: 8, 16: 8, 17: 8, 18: 8}
_TYPE_FMT = {1: "B", 2: "c", 3: "H", 4: "I", 6: "b", 7: "B", 8: "h", 9: "i", 11: "f", 12: "d", 16: "Q", 17: "q", 18: "Q"}
# numpy dtype <-> (sample format, bits)
_DTYPES = {np.dtype(np.uint8): (1, 8), np.dtype(np.uint16): (1, 16), np.dtype(np.uint32): (1, 32),
           np.dtype(np.int8): (2, 8), np.dtype(np.int16): (2, 16), np.dtype(np.int32): (2, 32), np.dtype(np.int64): (2, 64),
           np.dtype(np.uint64): (1, 64), np.dtype(np.float32): (3, 32), np.dtype(np.float64): (3, 64)}
_FROM_TIFF = {v: k for k, v in _DTYPES.items()}


# ------------------------------------------------------------------------------------------------ container: write
def _geo_tags(transform, crs):
    """ModelPixelScale / ModelTiepoint (or ModelTransformation for a rotated grid) + a minimal GeoKey directory."""
    tags = []
    if transform is not None:
        x0, sx, rx, y0, ry, sy = [float(v) for v in transform]
        if rx == 0.0 and ry == 0.0:
            tags.append((33550, _DOUBLE, [abs(sx), abs(sy), 0.0]))
            tags.append((33922, _DOUBLE, [0.0, 0.0, 0.0, x0, y0, 0.0]))
        else:
            tags.append((34264, _DOUBLE, [sx, rx, 0.0, x0, ry, sy, 0.0, y0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0]))
    cit = (str(crs) + "|") if crs else ""
    keys = [1, 1, 0, 0]
    entries = [(1024, 0, 1, 1 if crs else 32767), (1025, 0, 1, 1)]       # GTModelType (projected / user), RasterType area
    if cit:
        entries.append((1026, 34737, len(cit), 0))                       # GTCitationGeoKey -> GeoAsciiParams
    keys[3] = len(entries)
    for e in entries:
        keys.extend(e)
    tags.append((34735, _SHORT, keys))
    if cit:
        tags.append((34737, _ASCII, cit))
    return tags


def _pack_value(typ, values):
    if typ == _ASCII:
        b = values.encode("latin-1", "replace") + b"\0"
        return b, len(b)
    vals = list(values)
    return struct.pack("<%d%s" % (len(vals), _TYPE_FMT[typ]), *vals), len(vals)


def build_tiff(rows, cols, dtype, predictor, tile_sizes, transform=None, crs=None, nodata=None):
    """Header and tag directory of a tiled, deflate-compressed single-band TIFF whose tile data (the zlib streams, in
    tile order, one after the other) follows the header directly.  Returns (head bytes, tail bytes): the file is
    head + data + tail.  Classic TIFF below 4 GB, BigTIFF above (the reference's bigtiff='if_safer')."""
    fmt, bits = _DTYPES[np.dtype(dtype)]
    total = int(sum(int(s) for s in tile_sizes))
    big = total + 65536 + 16 * len(tile_sizes) >= (1 << 32) - (1 << 20)
    head_len = 16 if big else 8
    offsets, pos = [], head_len
    for s in tile_sizes:
        offsets.append(pos)
        pos += int(s)
    data_end = pos + (pos & 1)
    off_type = _LONG8 if big else _LONG
    tags = [(256, _LONG, [cols]), (257, _LONG, [rows]), (258, _SHORT, [bits]), (259, _SHORT, [8]), (262, _SHORT, [1]),
            (277, _SHORT, [1]), (284, _SHORT, [1]), (317, _SHORT, [predictor]), (322, _SHORT, [TILE]), (323, _SHORT, [TILE]),
            (324, off_type, offsets), (325, off_type, [int(s) for s in tile_sizes]), (339, _SHORT, [fmt])]
    tags += _geo_tags(transform, crs)
    if nodata is not None:
        tags.append((42113, _ASCII, "%.17g" % float(nodata)))
    tags.sort(key=lambda t: t[0])
    # the directory sits after the data; values that do not fit the entry go after the directory
    entry = 20 if big else 12
    ifd_off = data_end
    ifd_len = (8 if big else 2) + entry * len(tags) + (8 if big else 4)
    extra_off = ifd_off + ifd_len
    ifd = struct.pack("<Q" if big else "<H", len(tags))
    extra = b""
    inline = 8 if big else 4
    for tag, typ, values in tags:
        raw, count = _pack_value(typ, values)
        if len(raw) <= inline:
            field = raw + b"\0" * (inline - len(raw))
        else:
            if (extra_off + len(extra)) & 1:
                extra += b"\0"
            field = struct.pack("<Q" if big else "<I", extra_off + len(extra))
            extra += raw
        ifd += struct.pack("<HH", tag, typ) + struct.pack("<Q" if big else "<I", count) + field
    ifd += struct.pack("<Q" if big else "<I", 0)
    head = (b"II" + struct.pack("<HHHQ", 43, 8, 0, ifd_off)) if big else (b"II" + struct.pack("<HI", 42, ifd_off))
    tail = (b"\0" if (pos & 1) else b"") + ifd + extra
    return head, tail


# ------------------------------------------------------------------------------------------------- container: read
class TiffInfo(object):
    """The tags of the first image of a (Big)TIFF file that matter for a single-band raster."""

    def __init__(self, filepath):
        self.filepath = filepath
        with open(filepath, "rb") as f:
            head = f.read(16)
            if head[:2] == b"II":
                bo = "<"
            elif head[:2] == b"MM":
                bo = ">"
            else:
                raise ValueError("%s: not a TIFF file" % filepath)
            magic = struct.unpack(bo + "H", head[2:4])[0]
            if magic == 42:
                big, ifd_off = False, struct.unpack(bo + "I", head[4:8])[0]
            elif magic == 43:
                big, ifd_off = True, struct.unpack(bo + "Q", head[8:16])[0]
            else:
                raise ValueError("%s: not a TIFF file (magic %d)" % (filepath, magic))
            self.bo, self.big = bo, big
            f.seek(ifd_off)
            n = struct.unpack(bo + ("Q" if big else "H"), f.read(8 if big else 2))[0]
            entry = 20 if big else 12
            raw = f.read(entry * n)
            self.tags = {}
            for k in range(n):
                e = raw[k * entry:(k + 1) * entry]
                tag, typ = struct.unpack(bo + "HH", e[:4])
                count = struct.unpack(bo + ("Q" if big else "I"), e[4:12] if big else e[4:8])[0]
                field = e[12:20] if big else e[8:12]
                size = _TYPE_SIZE.get(typ, 1) * count
                if size <= len(field):
                    data = field[:size]
                else:
                    off = struct.unpack(bo + ("Q" if big else "I"), field)[0]
                    here = f.tell()
                    f.seek(off)
                    data = f.read(size)
                    f.seek(here)
                if typ == _ASCII:
                    self.tags[tag] = data.split(b"\0")[0].decode("latin-1")
                elif typ in _TYPE_FMT:
                    self.tags[tag] = list(struct.unpack(bo + "%d%s" % (count, _TYPE_FMT[typ]), data))
        t = self.tags
        self.cols, self.rows = int(t[256][0]), int(t[257][0])
        if int(t.get(277, [1])[0]) != 1:
            raise NotImplementedError("%s: only single-band rasters" % filepath)
        bits, fmt = int(t.get(258, [1])[0]), int(t.get(339, [1])[0])
        if (fmt, bits) not in _FROM_TIFF:
            raise NotImplementedError("%s: sample format %d with %d bits" % (filepath, fmt, bits))
        self.dtype, self.sample_format, self.sample_bytes = _FROM_TIFF[(fmt, bits)], fmt, bits // 8
        self.compression = int(t.get(259, [1])[0])
        if self.compression not in (1, 8, 32946):
            raise NotImplementedError("%s: compression %d (none and deflate are supported)" % (filepath, self.compression))
        self.predictor = int(t.get(317, [1])[0])
        if self.predictor not in (1, 2):
            raise NotImplementedError("%s: predictor %d" % (filepath, self.predictor))
        if bo == ">" and self.sample_bytes > 1:
            raise NotImplementedError("%s: big-endian samples" % filepath)
        if 322 in t:
            self.block_w, self.block_h = int(t[322][0]), int(t[323][0])
            self.offsets, self.counts = [int(v) for v in t[324]], [int(v) for v in t[325]]
        else:
            self.block_w = self.cols
            self.block_h = min(int(t.get(278, [self.rows])[0]), self.rows)
            self.offsets, self.counts = [int(v) for v in t[273]], [int(v) for v in t[279]]
        nd = t.get(42113)
        try:
            self.nodata = float(nd) if nd is not None and nd.strip() != "" else None
        except ValueError:
            self.nodata = None
        # GDAL style affine transform
        if 34264 in t:
            m = t[34264]
            self.transform = (m[3], m[0], m[1], m[7], m[4], m[5])
        elif 33550 in t and 33922 in t:
            sx, sy = t[33550][0], t[33550][1]
            i, j, _, x, y, _ = t[33922][:6]
            self.transform = (x - i * sx, sx, 0.0, y + j * sy, 0.0, -sy)
        else:
            self.transform = (0.0, 1.0, 0.0, 0.0, 0.0, 1.0)
        cit = t.get(34737)
        self.crs = cit.rstrip("|") if cit else ""


# ------------------------------------------------------------------------------------------------------- the classes
def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("malstroem_b200.io needs a CUDA device (there is no CPU fallback)")
    return torch


class RasterReader(object):
    """io.RasterReader (io.py:21-72): a single-band GeoTIFF into a 2-D numpy array, nodata replaced by
    `nodatasubst`.  read_device() returns the raster as a cuda tensor (for the device-resident pipeline)."""

    def __init__(self, filepath, nodatasubst=None):
        self.filepath = filepath
        self._info = TiffInfo(filepath)
        self.transform = self._info.transform
        self.crs = self._info.crs
        self.nodata = self._info.nodata
        self.nodatasubst = nodatasubst

    def read_device(self, device=0):
        torch = _torch()
        t = self._info
        dev = torch.device("cuda", device)
        nblocks = len(t.offsets)
        # the compressed blocks, one after the other, in one pinned buffer
        total = sum(t.counts)
        host = torch.empty(max(total, 1), dtype=torch.uint8).pin_memory()
        hv = host.numpy()
        offs = np.zeros(nblocks, np.uint64)
        pos = 0
        with open(self.filepath, "rb") as f:
            for k in range(nblocks):
                f.seek(t.offsets[k])
                f.readinto(memoryview(hv)[pos:pos + t.counts[k]])
                offs[k] = pos
                pos += t.counts[k]
        block_bytes = t.block_w * t.block_h * t.sample_bytes
        compressed = t.compression != 1
        d_in = host.to(dev, non_blocking=True)
        d_off = torch.from_numpy(offs.view(np.int64)).to(dev)
        d_len = torch.from_numpy(np.asarray(t.counts, dtype=np.uint32).view(np.int32)).to(dev)
        scratch = torch.empty(nblocks * block_bytes, dtype=torch.uint8, device=dev)
        if not compressed:
            # raw blocks: lined up at the block stride (the last strip may be short)
            for k in range(nblocks):
                o = int(offs[k])
                scratch[k * block_bytes:k * block_bytes + t.counts[k]] = d_in[o:o + t.counts[k]]
        out = torch.empty((t.rows, t.cols), dtype=getattr(torch, np.dtype(t.dtype).name), device=dev)
        # `if self.nodata and self.nodatasubst is not None` (io.py:69): a nodata value of 0 is never substituted
        mode = 0
        if self.nodata and self.nodatasubst is not None:
            mode = 2 if np.isnan(self.nodata) else 1
        stream = torch.cuda.current_stream(dev).cuda_stream
        with _lib.lock:
            _lib.check(_lib.lib().ms_init(dev.index or 0), "ms_init")
            _lib.check(_lib.lib().ms_tiff_decode_dev(
                d_in.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), nblocks, t.block_w, t.block_h, t.sample_bytes,
                t.sample_format, t.predictor, scratch.data_ptr(), out.data_ptr(), t.rows, t.cols, mode,
                float(self.nodata) if self.nodata is not None else 0.0,
                float(self.nodatasubst) if self.nodatasubst is not None else 0.0, 1 if compressed else 0,
                ctypes.c_void_p(stream)), "RasterReader.read")
        return out

    def read(self):
        return self.read_device().cpu().numpy()


class RasterWriter(object):
    """io.RasterWriter (io.py:75-139): a 2-D array (numpy, or a cuda tensor) as a tiled, deflate-compressed GeoTIFF,
    predictor 2 for float32 / int32 / uint8 and none for float64, as the reference configures GDAL."""

    def __init__(self, filepath, transform, crs, nodata=None):
        self.filepath = filepath
        self.transform = transform
        self.crs = crs
        self.driver = 'gtiff'
        self.options = dict(tiled='yes', compress='deflate', bigtiff='if_safer')
        self.datatype = None
        self.nodata = nodata
        self.stats = {}

    def write(self, data, device=0):
        torch = _torch()
        if isinstance(data, np.ndarray):
            dtype = data.dtype
        else:
            dtype = np.dtype(str(data.dtype).replace("torch.", ""))
        if not self.datatype:
            # io.py:124-136: the four types the reference knows, predictor 2 for all but float64
            if dtype == np.float64:
                self.datatype = 'Float64'
            elif dtype == np.float32:
                self.datatype = 'Float32'
                self.options['predictor'] = 2
            elif dtype == np.int32:
                self.datatype = 'Int32'
                self.options['predictor'] = 2
            elif dtype == np.uint8:
                self.datatype = 'Byte'
                self.options['predictor'] = 2
            else:
                raise NotImplementedError("Cannot determine GDAL datatype for numpy datatype {}".format(dtype))
        if len(data.shape) != 2:
            raise ValueError("RasterWriter.write: a 2-D array is required")
        dev = torch.device("cuda", device)
        d = torch.from_numpy(np.ascontiguousarray(data)).to(dev) if isinstance(data, np.ndarray) else data.contiguous()
        rows, cols = int(d.shape[0]), int(d.shape[1])
        es = d.element_size()
        predictor = int(self.options.get('predictor', 1))
        L = _lib.lib()
        tiles = (-(-rows // TILE)) * (-(-cols // TILE))
        with _lib.lock:
            _lib.check(L.ms_init(dev.index or 0), "ms_init")
            slot = int(L.ms_tiff_tile_slot(es))
        slots = torch.empty(tiles * slot, dtype=torch.uint8, device=dev)
        sizes = torch.empty(tiles, dtype=torch.int32, device=dev)
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with _lib.lock:
            _lib.check(L.ms_tiff_encode_dev(d.data_ptr(), es, rows, cols, predictor, slots.data_ptr(), slot,
                                            sizes.data_ptr(), stream), "RasterWriter.write")
        offs = torch.cumsum(sizes.long(), 0) - sizes.long()
        total = int(offs[-1] + sizes[-1])
        packed = torch.empty(total, dtype=torch.uint8, device=dev)
        with _lib.lock:
            _lib.check(L.ms_tiff_pack_dev(slots.data_ptr(), slot, sizes.data_ptr(), offs.data_ptr(), tiles,
                                          packed.data_ptr(), stream), "RasterWriter.write")
        host = torch.empty(total, dtype=torch.uint8).pin_memory()
        host.copy_(packed, non_blocking=True)
        tile_sizes = sizes.cpu().numpy().astype(np.int64)
        torch.cuda.current_stream(dev).synchronize()
        head, tail = build_tiff(rows, cols, dtype, predictor, tile_sizes, self.transform, self.crs, self.nodata)
        with open(self.filepath, "wb") as f:
            f.write(head)
            f.write(memoryview(host.numpy()))
            f.write(tail)
        self.stats = {"raw_bytes": rows * cols * es, "file_bytes": len(head) + total + len(tail), "tiles": tiles}
