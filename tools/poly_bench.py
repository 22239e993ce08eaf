"""Polygonisation timing: bluespot and watershed labels of a synthetic run, device resident.
usage: python tools/poly_bench.py [size] [repeats]"""
import sys
import time

import torch

sys.path.insert(0, ".")
from malstroem_b200 import _lib, pipeline, synth, vector  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    rep = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    p = pipeline.RasterPipeline(n, n)
    with _lib.lock:
        _lib.check(_lib.lib().ms_synth_fractal_dev(p.dem.data_ptr(), n, n, 0, 0, 1, None), "synth")
    p.run()
    torch.cuda.synchronize()
    for which in ("labels", "wsheds"):
        t = p.out[which]
        best = None
        for _ in range(rep):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rings = vector.polygonize_labels_device(t)
            dt = time.perf_counter() - t0
            best = dt if best is None or dt < best else best
        t0 = time.perf_counter()
        nfeat = sum(1 for _ in rings.features((0.0, 0.4, 0.0, 0.0, 0.0, -0.4)))
        tf = time.perf_counter() - t0
        print("%s %dx%d: %d rings (%d holes), %d vertices, %d boundary edges; device + fetch %.1f ms (%.2f Gcell/s); "
              "%d GeoJSON features built on the host in %.1f s"
              % (which, n, n, len(rings), int(rings.hole.sum()), len(rings.vrow), rings.nedges, best * 1e3,
                 n * n / best / 1e9, nfeat, tf))


if __name__ == "__main__":
    main()
