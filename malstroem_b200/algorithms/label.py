"""Mirror of malstroem/algorithms/label.py."""
import numpy as np

from .. import _lib

STATS_DTYPE = [('min', np.float64), ('max', np.float64), ('sum', np.float64), ('count', np.int64)]   # label.py:60
INDEX_DTYPE = [('value', np.float64), ('row', np.int64), ('col', np.int64)]                          # label.py:118


def _labels32(labelled, what):
    lab = np.asarray(labelled)
    if not (np.issubdtype(lab.dtype, np.integer) or lab.dtype == np.bool_):
        raise ValueError("%s: labelled must be an integer raster" % what)
    if lab.dtype != np.int32:
        if lab.size and (lab.max() > np.iinfo(np.int32).max or lab.min() < np.iinfo(np.int32).min):
            raise ValueError("%s: labels do not fit int32" % what)
        lab = lab.astype(np.int32)
    return np.ascontiguousarray(lab)


def _nlabels(lab, nlabels, what):
    # `if not nlabels: nlabels = np.max(labelled)` (label.py:56-57, 115-116, 149-150)
    if not nlabels:
        hi = np.zeros(1, np.int32)
        lo = np.zeros(1, np.int32)
        with _lib.lock:
            _lib.check(_lib.lib().ms_label_range(_lib.ptr(lab), lab.size, _lib.ptr(lo), _lib.ptr(hi)), what)
        _lib.track(lab)
        nlabels = int(hi[0])
    nlabels = int(nlabels)
    if nlabels < 0:
        raise ValueError("%s: negative label count" % what)
    return nlabels


def connected_components(data):
    """label.connected_components (label.py:19-40): scipy.ndimage.label with the full 3x3 structure —
    int32 labels numbered by first cell in row-major order, and the number of components."""
    data = np.asarray(data)
    if data.ndim != 2 or data.size == 0:
        raise ValueError("connected_components: a non-empty 2-D array is required")
    if data.dtype not in _lib.DTYPE_CODE:
        data = data != 0                 # any other dtype: the foreground mask is all scipy looks at
    d = np.ascontiguousarray(data)
    out = _lib.result_array(d.shape, np.int32)
    n = np.zeros(1, np.int64)
    with _lib.lock:
        _lib.check(_lib.lib().ms_connected_components(_lib.ptr(d.view(np.uint8) if d.dtype == np.bool_ else d),
                                                      _lib.DTYPE_CODE[d.dtype], _lib.ptr(out), d.shape[0],
                                                      d.shape[1], _lib.ptr(n)), "connected_components")
    _lib.track(d, out)
    return out, int(n[0])


def label_stats(data, labelled, nlabels=None):
    """label.label_stats (label.py:43-75): per label (0..nlabels) min, max, sum, count."""
    data = np.asarray(data)
    lab = _labels32(labelled, "label_stats")
    if data.shape != lab.shape or data.size == 0:
        raise ValueError("label_stats: data and labelled must have the same non-empty shape")
    if data.dtype not in (np.float32, np.float64):
        data = data.astype(np.float64)   # the reference's fallback converts every value to float64 (_label.pyx:49)
    d = np.ascontiguousarray(data)
    nlabels = _nlabels(lab, nlabels, "label_stats")
    stats = np.zeros((nlabels + 1,), dtype=STATS_DTYPE)
    mn, mx, sm = (np.empty(nlabels + 1, np.float64) for _ in range(3))
    cnt = np.empty(nlabels + 1, np.int64)
    with _lib.lock:
        _lib.check(_lib.lib().ms_label_stats(_lib.ptr(d), _lib.DTYPE_CODE[d.dtype], _lib.ptr(lab), d.size, nlabels,
                                             _lib.ptr(mn), _lib.ptr(mx), _lib.ptr(sm), _lib.ptr(cnt)), "label_stats")
    _lib.track(d, lab)
    stats['min'], stats['max'], stats['sum'], stats['count'] = mn, mx, sm, cnt
    return stats


def keep_labels(labelled, keep_label, background=0):
    """label.keep_labels (label.py:78-98).  Mutates keep_label[background] = False like the reference."""
    keep_label[background] = False
    keep = np.ascontiguousarray(np.array(keep_label).astype(bool)).view(np.uint8)
    lab = _labels32(labelled, "keep_labels")
    if lab.size == 0:
        return np.zeros(lab.shape, bool)
    out = _lib.result_array(lab.shape, np.uint8)
    with _lib.lock:
        _lib.check(_lib.lib().ms_keep_labels(_lib.ptr(lab), lab.size, _lib.ptr(keep), keep.size, _lib.ptr(out)),
                   "keep_labels")
    _lib.track(lab, out)
    return out.view(np.bool_)


def _extreme(data, labelled, nlabels, want_max, what):
    data = np.asarray(data)
    lab = _labels32(labelled, what)
    if data.ndim != 2 or data.shape != lab.shape or data.size == 0:
        raise ValueError("%s: data and labelled must be 2-D with the same non-empty shape" % what)
    d = np.ascontiguousarray(data, dtype=np.float64)     # record value is float64 (label.py:118)
    nlabels = _nlabels(lab, nlabels, what)
    out = np.zeros((nlabels + 1,), dtype=INDEX_DTYPE)
    val = np.empty(nlabels + 1, np.float64)
    row = np.empty(nlabels + 1, np.int64)
    col = np.empty(nlabels + 1, np.int64)
    with _lib.lock:
        _lib.check(_lib.lib().ms_label_extreme_index(_lib.ptr(d), _lib.ptr(lab), d.shape[0], d.shape[1], nlabels,
                                                     1 if want_max else 0, _lib.ptr(val), _lib.ptr(row),
                                                     _lib.ptr(col)), what)
    _lib.track(d, lab)
    out['value'], out['row'], out['col'] = val, row, col
    return out


def label_min_index(data, labelled, nlabels=None):
    """label.label_min_index (label.py:101-132): per label the minimum value and its first (row, col)."""
    return _extreme(data, labelled, nlabels, False, "label_min_index")


def label_max_index(data, labelled, nlabels=None):
    """label.label_max_index (label.py:135-166): per label the maximum value and its first (row, col)."""
    return _extreme(data, labelled, nlabels, True, "label_max_index")


def label_count(labelled):
    """label.label_count (label.py:169-180): np.bincount(labelled.ravel())."""
    lab = _labels32(labelled, "label_count")
    if lab.size == 0:
        return np.zeros(0, np.int64)
    lo, hi = np.zeros(1, np.int32), np.zeros(1, np.int32)
    with _lib.lock:
        _lib.check(_lib.lib().ms_label_range(_lib.ptr(lab), lab.size, _lib.ptr(lo), _lib.ptr(hi)), "label_count")
        if lo[0] < 0:
            raise ValueError("'list' argument must have no negative elements")      # np.bincount's message
        nb = int(hi[0]) + 1
        out = np.empty(nb, np.int64)
        _lib.check(_lib.lib().ms_label_count(_lib.ptr(lab), lab.size, nb, _lib.ptr(out)), "label_count")
    _lib.track(lab)
    return out
