"""Label rasters for the polygonisation tests (shared by the CPU oracle tests and the GPU parity tests)."""
import numpy as np


def cases():
    out = {}
    rng = np.random.default_rng(7)
    out["one_cell"] = np.array([[5]], dtype=np.int32)
    out["row"] = np.array([[1, 1, 2, 2, 2, 1, 0]], dtype=np.int32)
    out["col"] = np.array([[1], [1], [0], [3]], dtype=np.int32)
    out["uniform"] = np.full((9, 13), 4, dtype=np.int32)
    cb = (np.add.outer(np.arange(12), np.arange(17)) % 2).astype(np.int32)      # every interior vertex is a saddle
    out["checkerboard"] = cb
    nest = np.zeros((21, 21), dtype=np.int32)                                   # rings inside rings (holes in holes)
    for k, v in enumerate((1, 0, 2, 0, 1, 3)):
        nest[k * 2:21 - k * 2, k * 2:21 - k * 2] = v
    out["nested"] = nest
    diag = np.zeros((16, 16), dtype=np.int32)                                   # a region held together by corners only
    for k in range(16):
        diag[k, k] = 7
        diag[k, 15 - k] = 7
    out["diagonals"] = diag
    for k, (sh, nv) in enumerate((((5, 7), 2), ((17, 23), 3), ((40, 33), 2), ((64, 64), 5), ((97, 130), 3))):
        out["random%d" % k] = rng.integers(0, nv, size=sh).astype(np.int32)
    blobs = rng.random((70, 90))
    for _ in range(3):
        blobs = (blobs + np.roll(blobs, 1, 0) + np.roll(blobs, -1, 0) + np.roll(blobs, 1, 1) + np.roll(blobs, -1, 1)) / 5
    out["blobs"] = (np.digitize(blobs, np.quantile(blobs, [0.3, 0.6, 0.8]))).astype(np.int32)
    out["negative_values"] = (rng.integers(0, 3, size=(20, 20)) - 1).astype(np.int32) * 1000000
    return out
