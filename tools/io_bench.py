"""Throughput of the device GeoTIFF codec on the rasters of one run at S x S: per raster the device
time of encode + pack (CUDA events), the file size against the raw size, the wall time of RasterWriter.write including
the D2H copy and the file write (to /dev/shm), and the device time / wall time of reading it back.
usage: python tools/io_bench.py [S]"""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from malstroem_b200 import io as mio
from malstroem_b200.pipeline import RasterPipeline, synth_fractal

S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
p = RasterPipeline(S, S)
synth_fractal(S, S, seed=1, out=p.dem)
p.run()
torch.cuda.synchronize()
d = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
rasters = [("dem", p.dem)] + [(k, p.out[k]) for k in ("filled", "depths", "flowdir", "accum", "labels", "wsheds")]
tot_raw = tot_file = 0
tot_w = tot_r = 0.0
print("%-8s %-8s %9s %9s %7s %10s %10s %10s" % ("raster", "dtype", "raw MB", "file MB", "ratio", "write ms", "read ms", "GB/s w/r"))
for name, t in rasters:
    path = os.path.join(d, name + ".tif")
    w = mio.RasterWriter(path, (0.0, 0.4, 0.0, 0.0, 0.0, -0.4), "EPSG:25832")
    w.write(t)                      # warm-up (allocations)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    w.write(t)
    torch.cuda.synchronize(); tw = time.perf_counter() - t0
    r = mio.RasterReader(path)
    back = r.read_device()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    back = r.read_device()
    torch.cuda.synchronize(); tr = time.perf_counter() - t0
    assert torch.equal(back, t), name
    raw, fil = w.stats["raw_bytes"], w.stats["file_bytes"]
    tot_raw += raw; tot_file += fil; tot_w += tw; tot_r += tr
    print("%-8s %-8s %9.1f %9.1f %7.2f %10.1f %10.1f %5.1f/%4.1f" % (name, str(t.dtype).replace("torch.", ""), raw / 1e6, fil / 1e6,
          raw / fil, tw * 1e3, tr * 1e3, raw / tw / 1e9, raw / tr / 1e9))
    os.remove(path)
print("all: %.0f MB raw -> %.0f MB in files; write %.0f ms, read back %.0f ms (device codec + D2H / H2D + file I/O on %s)"
      % (tot_raw / 1e6, tot_file / 1e6, tot_w * 1e3, tot_r * 1e3, d))
