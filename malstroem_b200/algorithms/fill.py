"""Mirror of malstroem/algorithms/fill.py (same names, arguments and error behaviour), computed by the
sm_100a library through the C ABI (include/malstroem_b200.h)."""
import numpy as np

from .. import _lib

DTYPE_DTM = np.float32            # malstroem/algorithms/dtypes.py:20-31
DTYPE_FILL = np.float32
DTYPE_FILLNOFLAT = np.float64


def _dtm2d(dtm, what):
    dtm = np.asarray(dtm)
    if dtm.ndim != 2:
        raise ValueError("%s: Buffer has wrong number of dimensions (expected 2, got %d)" % (what, dtm.ndim))
    if dtm.dtype != np.float32:
        # the compiled reference sweep only accepts float32 (speedups/_fill.pyx:30, "Buffer dtype mismatch")
        raise ValueError("%s: Buffer dtype mismatch, expected 'float32' but got '%s'" % (what, dtm.dtype))
    if dtm.shape[0] <= 3 or dtm.shape[1] <= 3:
        # speedups/_fill.pyx:31-32 (3 rows/cols); fewer is undefined in the reference
        raise ValueError("Width or height of processing area is zero")
    return np.ascontiguousarray(dtm)


def fill_terrain(dtm):
    """fill.fill_terrain (fill.py:112-171): depressionless float32 DEM, same shape."""
    dtm = _dtm2d(dtm, "fill_terrain")
    out = _lib.result_array(dtm.shape, DTYPE_FILL)
    with _lib.lock:
        _lib.check(_lib.lib().ms_fill_terrain(_lib.ptr(dtm), _lib.ptr(out), None, dtm.shape[0], dtm.shape[1]),
                   "fill_terrain")
    _lib.track(dtm, out)
    return out


def fill_terrain_and_depths(dtm):
    """fill_terrain plus `filled - dtm` (dem.py:67-71) in one device pass."""
    dtm = _dtm2d(dtm, "fill_terrain")
    out = _lib.result_array(dtm.shape, DTYPE_FILL)
    dep = _lib.result_array(dtm.shape, DTYPE_FILL)
    with _lib.lock:
        _lib.check(_lib.lib().ms_fill_terrain(_lib.ptr(dtm), _lib.ptr(out), _lib.ptr(dep), dtm.shape[0], dtm.shape[1]),
                   "fill_terrain")
    _lib.track(dtm, out, dep)
    return out, dep


def fill_terrain_no_flats(dtm, short=0, diag=0):
    """fill.fill_terrain_no_flats (fill.py:174-232): float64 surface with a strictly descending path."""
    dtm = _dtm2d(dtm, "fill_terrain_no_flats")
    out = _lib.result_array(dtm.shape, DTYPE_FILLNOFLAT)
    with _lib.lock:
        _lib.check(_lib.lib().ms_fill_terrain_no_flats(_lib.ptr(dtm), float(short), float(diag), _lib.ptr(out),
                                                       dtm.shape[0], dtm.shape[1]), "fill_terrain_no_flats")
    _lib.track(dtm, out)
    return out


def minimum_safe_short_and_diag(dem):
    """fill.minimum_safe_short_and_diag (fill.py:235-250).  The raster pass (min / max) runs on the device;
    the two scalar operations follow the reference literally."""
    dem = np.asarray(dem)
    if dem.dtype == np.float32 and dem.size:
        d = np.ascontiguousarray(dem)
        lo, hi = np.zeros(1, np.float32), np.zeros(1, np.float32)
        with _lib.lock:
            _lib.check(_lib.lib().ms_minmax_f32(_lib.ptr(d), d.size, _lib.ptr(lo), _lib.ptr(hi)),
                       "minimum_safe_short_and_diag")
        _lib.track(d)
        amax, amin = hi[0], lo[0]
    else:
        amax, amin = np.amax(dem), np.amin(dem)
    maxval = DTYPE_FILLNOFLAT(max(abs(amax), abs(amin)))
    nextval = np.nextafter(maxval, DTYPE_FILLNOFLAT(float('inf')))
    short = (nextval - maxval) * 1024
    diag = short * (2 ** 0.5)
    return short, diag
