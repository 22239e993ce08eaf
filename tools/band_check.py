"""Row-band decomposition on ONE GPU (bands as threads) against the single-GPU path: bit-exact comparison."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from malstroem_b200.pipeline import RasterPipeline, synth_fractal
from malstroem_b200 import bands

R = int(sys.argv[1]) if len(sys.argv) > 1 else 512
C = int(sys.argv[2]) if len(sys.argv) > 2 else 384
Gs = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [2, 3]
seed = int(sys.argv[4]) if len(sys.argv) > 4 else 1
dem = synth_fractal(R, C, seed=seed)
ref = RasterPipeline(R, C).run(dem)
torch.cuda.synchronize()
ok_all = True
for G in Gs:
    t0 = time.perf_counter()
    pipes = bands.run_threaded(dem, G)
    dt = time.perf_counter() - t0
    msgs = []
    for name in ("filled", "depths", "fnf", "flowdir", "accum", "labels", "wsheds"):
        got = torch.cat([p.out[name] for p in pipes])
        same = torch.equal(got, ref.out[name])
        if not same:
            bad = (got != ref.out[name])
            msgs.append("%s: %d cells differ (first at %s)" % (name, int(bad.sum()), bad.nonzero()[0].tolist()))
    p0 = pipes[0]
    if p0.nlabels != ref.nlabels:
        msgs.append("nlabels %d != %d" % (p0.nlabels, ref.nlabels))
    else:
        m = ref.nlabels + 1
        for name in ("st_min", "st_max", "st_count", "ws_count", "ppmin_row", "ppmin_col", "ppmax_row", "ppmax_col", "ppmin_value", "ppmax_value"):
            if not torch.equal(p0.tables[name][:m], ref.tables[name][:m]):
                d = (p0.tables[name][:m] != ref.tables[name][:m])
                msgs.append("table %s: %d entries differ (first %s)" % (name, int(d.sum()), d.nonzero()[0].tolist()))
        if not torch.allclose(p0.tables["st_sum"][:m], ref.tables["st_sum"][:m], rtol=1e-6, atol=0):
            msgs.append("table st_sum differs beyond 1e-6")
    print("G=%d %dx%d: %s  (%.1f ms)  stats %s" % (G, R, C, "OK" if not msgs else "MISMATCH", dt * 1e3, p0.stats))
    for m_ in msgs:
        print("   ", m_)
    ok_all &= not msgs
    for p in pipes:
        p.close()
sys.exit(0 if ok_all else 1)
