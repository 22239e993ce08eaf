"""Dump the per-visit log of the tail of k_nf_solve_ir (library built with -DNF_STATS -DNF_STATS_TAIL, MS_LIB pointing
at it) to gpurun_out/nf_tail_<S>.npy: columns pop, loaded, end [ns], tile, rounds, relax [ns].
usage: python tools/nf_chain.py [S] ; analysis: python tools/nf_chain.py --analyse gpurun_out/nf_tail_<S>.npy <tiles_x>"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def analyse(path, tiles_x):
    a = np.load(path)
    pop, loaded, end, tile, rounds, relax = (a[:, k] for k in range(6))
    order = np.argsort(pop)
    pop, loaded, end, tile, rounds, relax = (x[order] for x in (pop, loaded, end, tile, rounds, relax))
    t0 = pop.min()
    n = len(pop)
    print("tail visits %d over %.2f ms; distinct tiles %d; visits per tile max %d" % (
        n, (end.max() - t0) / 1e6, len(np.unique(tile)), np.bincount(np.unique(tile, return_inverse=True)[1]).max()))
    ty, tx = tile // tiles_x, tile % tiles_x
    # predecessor of a visit: the latest earlier visit on a neighbouring (or the same) tile that was loaded before this pop
    pred = np.full(n, -1)
    depth = np.zeros(n, dtype=np.int64)
    for v in range(n):
        lo = max(0, v - 4000)
        cand = np.flatnonzero((np.abs(ty[lo:v] - ty[v]) <= 1) & (np.abs(tx[lo:v] - tx[v]) <= 1) & (loaded[lo:v] < pop[v])) + lo
        if len(cand):
            u = cand[np.argmax(loaded[cand])]
            pred[v] = u
            depth[v] = depth[u] + 1
    last = int(np.argmax(end))
    chain = []
    v = last
    while v >= 0:
        chain.append(v)
        v = pred[v]
    chain = chain[::-1]
    print("chain ending at the last visit: %d hops, %.2f ms" % (len(chain), (end[last] - pop[chain[0]]) / 1e6))
    hop = np.diff(pop[chain]) / 1e3
    print("hop latency [us] (pop to pop): mean %.1f p50 %.1f p90 %.1f" % (hop.mean(), np.median(hop), np.percentile(hop, 90)))
    ct = tile[chain]
    print("distinct tiles on the chain %d; same-tile or back-and-forth hops %d" % (
        len(np.unique(ct)), int((ct[2:] == ct[:-2]).sum())))
    print("per chain visit [us]: load %.1f relax %.1f (rounds %.1f) flush %.1f" % (
        ((loaded - pop)[chain]).mean() / 1e3, relax[chain].mean() / 1e3, rounds[chain].mean(),
        ((end - loaded - relax)[chain]).mean() / 1e3))
    print("first 40 hops: tile (y,x) rounds  pop->pop us")
    for k in range(min(40, len(chain) - 1)):
        v = chain[k]
        print("  (%d,%d) r%d  %.1f" % (ty[v], tx[v], rounds[v], hop[k]))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--analyse":
        return analyse(sys.argv[2], int(sys.argv[3]))
    import ctypes
    import torch
    from malstroem_b200 import _lib
    from malstroem_b200.pipeline import synth_fractal
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    L = _lib.lib()
    L.ms_init(0)
    raw = ctypes.CDLL(_lib.LIB_PATH)
    dev = torch.device("cuda", 0)
    dem = synth_fractal(S, S, seed=1)
    filled = torch.empty_like(dem)
    depths = torch.empty_like(dem)
    fnf = torch.empty((S, S), dtype=torch.float64, device=dev)
    sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert L.ms_fill_terrain_dev(dem.data_ptr(), filled.data_ptr(), depths.data_ptr(), S, S, sp) == 0
    mv = np.float64(float(dem.abs().max()))
    sh = float((np.nextafter(mv, np.inf) - mv) * 1024)
    dg = sh * 2 ** 0.5
    nflog = raw.ms_nf_log
    nflog.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    for rep in range(3):
        nflog(None, None, 1)
        st = (ctypes.c_int64 * 8)()
        rc = L.ms_fill_terrain_no_flats_dev(dem.data_ptr(), filled.data_ptr(), sh, dg, fnf.data_ptr(), S, S, st, sp)
        assert rc == 0, L.ms_last_error()
        torch.cuda.synchronize()
    log = np.zeros(4 * 262144, dtype=np.uint64)
    nl = ctypes.c_uint(0)
    nflog(log.ctypes.data_as(ctypes.c_void_p), ctypes.byref(nl), 0)
    k = min(nl.value, 262144)
    log = log[: 4 * k].reshape(k, 4).astype(np.int64)
    out = np.stack([log[:, 0], log[:, 1], log[:, 2], log[:, 3] & 0xffffffff, (log[:, 3] >> 32) & 0xff,
                    (log[:, 3] >> 40) & 0xffffff], axis=1)
    os.makedirs("gpurun_out", exist_ok=True)
    np.save("gpurun_out/nf_tail_%d.npy" % S, out)
    print("visits %d, logged %d (tail mode) -> gpurun_out/nf_tail_%d.npy" % (st[1], k, S))


if __name__ == "__main__":
    main()
