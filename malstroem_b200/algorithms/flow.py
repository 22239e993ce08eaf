"""Mirror of malstroem/algorithms/flow.py for the three raster-wide functions of the hot path."""
import numpy as np

from .. import _lib

DTYPE_FLOWDIR = np.uint8          # malstroem/algorithms/dtypes.py:20-31
DTYPE_ACCUM = np.float64

# AGNPS direction codes (flow.py:30-38)
FLOWDIR_UP, FLOWDIR_UP_RIGHT, FLOWDIR_RIGHT, FLOWDIR_DOWN_RIGHT = 0, 1, 2, 3
FLOWDIR_DOWN, FLOWDIR_DOWN_LEFT, FLOWDIR_LEFT, FLOWDIR_UP_LEFT, FLOWDIR_NODIR = 4, 5, 6, 7, 8


def _arr2d(a, dtype, what):
    a = np.asarray(a)
    if a.ndim != 2:
        raise ValueError("%s: Buffer has wrong number of dimensions (expected 2, got %d)" % (what, a.ndim))
    if a.dtype != dtype:
        # compiled reference: typed memoryview argument (speedups/_flow.pyx:99,257)
        raise ValueError("%s: Buffer dtype mismatch, expected '%s' but got '%s'" % (what, np.dtype(dtype), a.dtype))
    if a.size == 0:
        raise ValueError("%s: empty raster" % what)
    return np.ascontiguousarray(a)


def terrain_flowdirection(terrain, edges_flow_outward=True):
    """flow.terrain_flowdirection (flow.py:142-167): uint8 D8 codes 0..7, 8 = no direction."""
    t = _arr2d(terrain, np.float64, "terrain_flowdirection")
    out = _lib.result_array(t.shape, DTYPE_FLOWDIR)
    with _lib.lock:
        _lib.check(_lib.lib().ms_flowdir(_lib.ptr(t), _lib.ptr(out), t.shape[0], t.shape[1],
                                         1 if edges_flow_outward else 0), "terrain_flowdirection")
    _lib.track(t, out)
    return out


def accumulated_flow(flowdir):
    """flow.accumulated_flow (flow.py:344-364): float64 upstream cell counts (including the cell)."""
    fd = _arr2d(flowdir, np.uint8, "accumulated_flow")
    out = _lib.result_array(fd.shape, DTYPE_ACCUM)
    with _lib.lock:
        _lib.check(_lib.lib().ms_accumulated_flow(_lib.ptr(fd), _lib.ptr(out), fd.shape[0], fd.shape[1]),
                   "accumulated_flow")
    _lib.track(fd, out)
    return out


def watersheds_from_labels(flowdir, labelled, unassigned):
    """flow.watersheds_from_labels (flow.py:398-412): IN PLACE, returns None.  int32 and int64 label
    rasters run natively (the reference's two typed variants, _flow.pyx:276-357); any other integer dtype
    goes through int64 like the reference's fallback variant (:360-403)."""
    fd = _arr2d(flowdir, np.uint8, "watersheds_from_labels")
    if not isinstance(labelled, np.ndarray) or labelled.shape != fd.shape:
        raise ValueError("watersheds_from_labels: labelled must be an ndarray of the flowdir's shape")
    if labelled.dtype in (np.int32, np.int64) and labelled.flags.c_contiguous:
        work = labelled
    elif np.issubdtype(labelled.dtype, np.integer):
        work = np.ascontiguousarray(labelled, dtype=np.int64 if labelled.dtype.itemsize > 4 or
                                    labelled.dtype == np.uint32 else np.int32)
    else:
        raise ValueError("watersheds_from_labels: labelled must be an integer raster")
    with _lib.lock:
        _lib.check(_lib.lib().ms_watersheds_from_labels(_lib.ptr(fd), _lib.ptr(work), work.dtype.itemsize,
                                                        fd.shape[0], fd.shape[1], int(unassigned)),
                   "watersheds_from_labels")
    _lib.track(fd, work)
    if work is not labelled:
        labelled[...] = work.astype(labelled.dtype)
    return None
