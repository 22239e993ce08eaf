"""CPU model of the order in which the no-flats tile solver visits its tiles (tools/nf_order_sim.c): the same lake mask
and seeds as K2 on the synthetic fractal DEM, a visit relaxes a tile to convergence, `P` tiles are in flight at once.
Compares the FIFO of k_nf_solve_ir with serving the smallest offered distance first (exact, and in buckets).
usage: python tools/nf_order_sim.py [size]   (result for 8192: DESIGN.md section 6)"""
import ctypes
import os
import subprocess
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from malstroem_b200 import synth
from oracle import port
S = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
dem = synth.fractal_dem(S, S, seed=1)
t0 = time.time(); F = port.fill_terrain(dem); print("fill %.1fs" % (time.time() - t0))
INF, WALL, SQ, DQ = 0x3fffffff, 0x7fffffff, 1024, 1448
Fp = np.pad(F, 1, constant_values=np.inf)
m = np.full(F.shape, np.inf, dtype=np.float32)
for dr in (-1, 0, 1):
    for dc in (-1, 0, 1):
        if dr or dc:
            m = np.minimum(m, Fp[1 + dr:1 + dr + S, 1 + dc:1 + dc + S])
fixed = (F == dem) & (m < F)
fixed[0] = fixed[-1] = True; fixed[:, 0] = fixed[:, -1] = True
D0 = np.where(fixed, WALL, INF).astype(np.int32)
# sources: non-fixed cells next to a fixed cell of the same level
fx = np.pad(fixed, 1, constant_values=False); Fq = np.pad(F, 1, constant_values=np.nan)
for dr in (-1, 0, 1):
    for dc in (-1, 0, 1):
        if dr or dc:
            nb_fixed = fx[1 + dr:1 + dr + S, 1 + dc:1 + dc + S]
            nb_F = Fq[1 + dr:1 + dr + S, 1 + dc:1 + dc + S]
            w = DQ if (dr and dc) else SQ
            src = (~fixed) & nb_fixed & (nb_F == F)
            D0 = np.where(src & (D0 > w), w, D0).astype(np.int32)
print("non-fixed %.1f%%  sources %d" % (100 * (~fixed).mean(), int(((D0 < INF)).sum())))
here = os.path.dirname(os.path.abspath(__file__))
so = '/tmp/nf_order_sim.so'
subprocess.run(['gcc', '-O2', '-shared', '-fPIC', '-o', so, os.path.join(here, 'nf_order_sim.c')], check=True)
L = ctypes.CDLL(so)
L.simulate.restype = ctypes.c_long
L.simulate.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
ntiles = (S // 64) ** 2
ref = None
for mode, shift, P, name in ((0, 0, 1, "FIFO sequential"), (0, 0, 888, "FIFO 888 in flight"), (1, 0, 888, "priority, 888 in flight"), (2, 17, 888, "buckets 2^17, 888 in flight")):
    D = D0.copy()
    sw, bt = ctypes.c_long(0), ctypes.c_long(0)
    t0 = time.time()
    v = L.simulate(D.ctypes.data, S, S, mode, shift, P, ctypes.byref(sw), ctypes.byref(bt))
    if ref is None: ref = D.copy()
    assert np.array_equal(ref, D)
    print("%-30s visits %7d (%.2f per tile)  in-tile sweeps %8d  batches %6d   %.1fs" % (name, v, v / ntiles, sw.value, bt.value, time.time() - t0), flush=True)
