"""CPU tests of the row-band host logic: band splitting, the boundary arithmetic of
malstroem_b200/bands.py (accumulation forest, watershed chains, component re-rooting plan), the library's host-side
boundary merge (ms_cc_boundary_merge — pure CPU code in the .so), and the torch.distributed exchange layer under
gloo with world_size 2.  The band-local kernel phases are CUDA (tests/test_bands_gpu.py); here they are emulated with
the CPU oracle on the band's rows, so that what is tested is exactly the part that runs on the host."""
import ctypes
import os
import socket

import numpy as np
import pytest
import torch

from malstroem_b200 import _lib, bands, synth
from oracle import port

DR = (-1, -1, 0, 1, 1, 1, 0, -1)
DC = (0, 1, 1, 1, 0, -1, -1, -1)


def surfaces(rows, cols, seed):
    dem = synth.fractal_dem(rows, cols, seed=seed)
    filled = port.fill_terrain(dem)
    short, diag = port.minimum_safe_short_and_diag(dem)
    fnf = port.fill_terrain_no_flats(dem, short, diag)
    fd = port.terrain_flowdirection(fnf)
    lab, n = port.connected_components(filled - dem)
    return dem, filled, fd, lab, n


def test_band_rows():
    assert bands.band_rows(8192, 8) == [(k * 1024, (k + 1) * 1024) for k in range(8)]
    assert bands.band_rows(320, 2) == [(0, 192), (192, 320)]
    r = bands.band_rows(1000, 3)
    assert r[0][0] == 0 and r[-1][1] == 1000 and all(a[1] == b[0] for a, b in zip(r, r[1:]))
    assert all((a[1] - a[0]) % 64 == 0 for a in r[:-1])
    with pytest.raises(ValueError):
        bands.band_rows(100, 3)


# ------------------------------------------------------------------------------ band-local phases, emulated
def band_trace(fd, r0, r1, r, c):
    """Follow the D8 path from (r, c) inside rows [r0, r1): returns ('exit', side, col, to_col) | ('end', r, c)."""
    rows, cols = fd.shape
    while True:
        d = fd[r, c]
        if d > 7:
            return ("end", r, c)
        nr, nc = r + DR[d], c + DC[d]
        if nc < 0 or nc >= cols or nr < 0 or nr >= rows:
            return ("end", r, c)
        if nr < r0:
            return ("exit", 0, c, nc)
        if nr >= r1:
            return ("exit", 1, c, nc)
        r, c = nr, nc


def accum_band_local(fd, r0, r1, top_open, bot_open):
    cols = fd.shape[1]
    a_loc = port.accumulated_flow(np.ascontiguousarray(fd[r0:r1]), fast=True)      # stepping out of the band ends a path
    exit_to = -np.ones(2 * cols, np.int32)
    exit_val = np.zeros(2 * cols, np.float64)
    entry_root = -np.ones(2 * cols, np.int32)
    for side, (is_open, r) in enumerate(((top_open, r0), (bot_open, r1 - 1))):
        if not is_open:
            continue
        for c in range(cols):
            d = fd[r, c]
            if d <= 7 and DR[d] == (1 if side else -1) and 0 <= c + DC[d] < cols:
                exit_to[side * cols + c] = c + DC[d]
                exit_val[side * cols + c] = a_loc[r - r0, c]
            t = band_trace(fd, r0 if top_open else -10**9, r1 if bot_open else 10**9, r, c)
            if t[0] == "exit":
                entry_root[side * cols + c] = t[1] * cols + t[2]
    return exit_to, exit_val, entry_root


def forest_totals(parent, vals):
    tot = vals.astype(np.float64).copy()
    indeg = np.zeros(len(parent), np.int64)
    for p in parent:
        if p >= 0:
            indeg[p] += 1
    stack = [i for i in range(len(parent)) if indeg[i] == 0]
    while stack:
        i = stack.pop()
        p = parent[i]
        if p >= 0:
            tot[p] += tot[i]
            indeg[p] -= 1
            if indeg[p] == 0:
                stack.append(p)
    return tot


@pytest.mark.parametrize("nb", [2, 3])
def test_accum_forest_against_oracle(nb):
    rows, cols = 192 * nb - 40, 160
    _, _, fd, _, _ = surfaces(rows, cols, seed=21)
    acc = port.accumulated_flow(fd, fast=True)
    rr = bands.band_rows(rows, nb)
    loc = [accum_band_local(fd, r0, r1, g > 0, g + 1 < nb) for g, (r0, r1) in enumerate(rr)]
    to = torch.from_numpy(np.stack([x[0] for x in loc]))
    root = torch.from_numpy(np.stack([x[2] for x in loc]))
    parent = bands.accum_forest(to, root, cols).numpy()
    tot = forest_totals(parent, np.concatenate([x[1] for x in loc])).reshape(nb, 2, cols)
    nexit = 0
    for g, (r0, r1) in enumerate(rr):
        for side, r in ((0, r0), (1, r1 - 1)):
            for c in range(cols):
                if loc[g][0][side * cols + c] >= 0:
                    nexit += 1
                    assert tot[g, side, c] == acc[r, c], (g, side, c)
    assert nexit > 20


def ws_band_local(fd, lab, r0, r1, top_open, bot_open):
    cols = fd.shape[1]
    edge_res = np.zeros(2 * cols, np.int32)
    exit_to = -np.ones(2 * cols, np.int32)
    lo, hi = (r0 if top_open else -10**9), (r1 if bot_open else 10**9)
    for side, (is_open, r) in enumerate(((top_open, r0), (bot_open, r1 - 1))):
        if not is_open:
            continue
        for c in range(cols):
            d = fd[r, c]
            if d <= 7 and DR[d] == (1 if side else -1) and 0 <= c + DC[d] < cols:
                exit_to[side * cols + c] = c + DC[d]
            # first labelled cell on the in-band path, else the exit, else nothing
            rr_, cc_ = r, c
            res = 0
            while True:
                if lab[rr_, cc_] != 0:
                    res = int(lab[rr_, cc_])
                    break
                dd = fd[rr_, cc_]
                if dd > 7:
                    break
                nr, nc = rr_ + DR[dd], cc_ + DC[dd]
                if nc < 0 or nc >= cols or nr < 0 or nr >= fd.shape[0]:
                    break
                if nr < lo or nr >= hi:
                    res = -(1 + (0 if nr < lo else 1) * cols + cc_)
                    break
                rr_, cc_ = nr, nc
            edge_res[side * cols + c] = res
    return edge_res, exit_to


@pytest.mark.parametrize("nb", [2, 3])
def test_watershed_chain_against_oracle(nb):
    rows, cols = 192 * nb - 70, 144
    _, _, fd, lab, _ = surfaces(rows, cols, seed=22)
    ws = lab.copy()
    port.watersheds_from_labels(fd, ws, 0)
    rr = bands.band_rows(rows, nb)
    loc = [ws_band_local(fd, lab, r0, r1, g > 0, g + 1 < nb) for g, (r0, r1) in enumerate(rr)]
    res = torch.from_numpy(np.stack([x[0] for x in loc]))
    to = torch.from_numpy(np.stack([x[1] for x in loc]))
    arr = bands.watershed_chain(res, to, cols).numpy()
    final = np.zeros_like(arr)
    for i, v in enumerate(arr):
        while v < 0:
            v = arr[-(v + 1)]
        final[i] = v
    final = final.reshape(nb, 2, cols)
    for g, (r0, r1) in enumerate(rr):
        for side, r in ((0, r0), (1, r1 - 1)):
            if (side == 0 and g == 0) or (side == 1 and g + 1 == nb):
                continue
            assert np.array_equal(final[g, side], ws[r]), (g, side)
        node, ok = bands.exit_targets(to, cols, g)
        node, ok = node.numpy(), ok.numpy()
        for side, r in ((0, r0), (1, r1 - 1)):
            for c in range(cols):
                k = side * cols + c
                if ok[k] and lab[r, c] == 0:
                    assert final.reshape(-1)[node[k]] == ws[r, c]


def cc_roots(lab_band, cell_offset):
    """Per cell the smallest (global) flat index of its band-local component, -1 for background."""
    flat = lab_band.ravel()
    first = np.full(flat.max() + 1, -1, np.int64)
    idx = np.arange(flat.size)[::-1]
    first[flat[::-1]] = idx                      # the last write wins = the smallest index
    out = np.where(flat > 0, first[flat] + cell_offset, -1)
    return out.reshape(lab_band.shape)


def merge_host(tops, bots):
    L = _lib.lib()
    G, cols = tops.shape
    cap = 2 * G * cols
    out_root, out_glob = np.empty(cap, np.int64), np.empty(cap, np.int64)
    k = ctypes.c_int64(0)
    top, bot = np.ascontiguousarray(tops), np.ascontiguousarray(bots)
    rc = L.ms_cc_boundary_merge(G, cols, _lib.ptr(top), _lib.ptr(bot), _lib.ptr(out_root), _lib.ptr(out_glob), cap,
                                ctypes.byref(k))
    assert rc == 0, L.ms_last_error()
    return out_root[: k.value], out_glob[: k.value]


def banded_labels(depths, nb):
    """The numbering steps of BandPipeline._labels with the kernel phases emulated in numpy."""
    rows, cols = depths.shape
    rr = bands.band_rows(rows, nb)
    roots = [cc_roots(port.connected_components(np.ascontiguousarray(depths[r0:r1]))[0], r0 * cols) for r0, r1 in rr]
    tops = np.stack([r[0] for r in roots])
    bots = np.stack([r[-1] for r in roots])
    mr, mg = merge_host(tops, bots)
    plans = [bands.cc_plan(mr, mg, r0 * cols, r1 * cols) for r0, r1 in rr]
    counts, ranks = [], []
    for (r0, r1), rt, pl in zip(rr, roots, plans):
        flat = rt.ravel() - r0 * cols
        isroot = (flat == np.arange(flat.size))
        isroot[pl["rerooted_local"]] = False
        ranks.append(np.cumsum(isroot) - isroot)          # exclusive scan
        counts.append(int(isroot.sum()))
    offs = np.concatenate([[0], np.cumsum(counts)[:-1]])
    comp_roots = plans[0]["comp_roots"]
    lab_of = np.zeros(len(comp_roots), np.int64)
    for g, pl in enumerate(plans):
        lab_of[pl["owned_pos"]] += ranks[g][pl["owned_local"]] + offs[g] + 1
    out = []
    for g, ((r0, r1), rt, pl) in enumerate(zip(rr, roots, plans)):
        flat = rt.ravel() - r0 * cols
        rk = ranks[g].astype(np.int64).copy()
        rk[pl["rerooted_local"]] = lab_of[np.searchsorted(comp_roots, pl["rerooted_global"])] - offs[g] - 1
        out.append(np.where(rt.ravel() >= 0, rk[np.maximum(flat, 0)] + offs[g] + 1, 0).reshape(rt.shape))
    return np.concatenate(out).astype(np.int32), int(sum(counts))


@pytest.mark.parametrize("nb,seed", [(2, 31), (3, 32), (4, 33)])
def test_cc_boundary_merge_and_numbering(nb, seed):
    rows, cols = 64 * 3 * nb - 17, 150
    dem, filled, _, lab, n = surfaces(rows, cols, seed)
    got, ngot = banded_labels(filled - dem, nb)
    assert ngot == n and np.array_equal(got, lab)


def test_cc_merge_snake_through_all_bands():
    """One component that crosses every band edge several times, and touches its neighbours only diagonally."""
    rows, cols, nb = 64 * 4, 40, 4
    d = np.zeros((rows, cols), np.float32)
    d[:, 30] = 1                                   # a vertical bar through all bands
    for k in range(1, nb):
        d[64 * k - 1, 5] = 1                       # diagonal contacts across the band edges
        d[64 * k, 6] = 1
        d[64 * k - 1, 6:30] = 1                    # ... joined to the bar only above the edge
    d[10, 2] = 1                                   # an isolated early component (takes label 1)
    lab, n = port.connected_components(d)
    got, ngot = banded_labels(d, nb)
    assert ngot == n and np.array_equal(got, lab)


# -------------------------------------------------------------------------------- gloo, world_size 2
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port_no, rows, cols, result_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = bands.DistComm()
        _, _, fd, lab, _ = surfaces(rows, cols, seed=41)
        acc = port.accumulated_flow(fd, fast=True)
        ws = lab.copy()
        port.watersheds_from_labels(fd, ws, 0)
        r0, r1 = bands.band_rows(rows, world)[rank]
        # halo exchange of a raster row pair
        up, down = comm.exchange(torch.from_numpy(fd[r0].copy()), torch.from_numpy(fd[r1 - 1].copy()))
        if rank > 0:
            assert np.array_equal(up.numpy(), fd[r0 - 1])
        else:
            assert up is None
        if rank + 1 < world:
            assert np.array_equal(down.numpy(), fd[r1])
        else:
            assert down is None
        # variable-length gather and reductions
        parts = comm.all_gather_var(torch.arange(3 + 2 * rank, dtype=torch.int32))
        assert [len(p) for p in parts] == [3 + 2 * k for k in range(world)]
        t = torch.tensor([float(rank + 1), -float(rank)], dtype=torch.float64)
        assert comm.all_reduce(t.clone(), "sum").tolist() == [3.0, -1.0]
        assert comm.all_reduce(t.clone(), "min").tolist() == [1.0, -1.0]
        # accumulation: band-local phase (oracle), exchange, forest, totals at this band's exits
        e_to, e_val, e_root = accum_band_local(fd, r0, r1, rank > 0, rank + 1 < world)
        a_to = comm.all_gather(torch.from_numpy(e_to))
        a_val = comm.all_gather(torch.from_numpy(e_val))
        a_root = comm.all_gather(torch.from_numpy(e_root))
        parent = bands.accum_forest(a_to, a_root, cols).numpy()
        tot = forest_totals(parent, a_val.numpy().reshape(-1)).reshape(world, 2, cols)
        for side, r in ((0, r0), (1, r1 - 1)):
            for c in range(cols):
                if e_to[side * cols + c] >= 0:
                    assert tot[rank, side, c] == acc[r, c]
        # watersheds: chains across the band edge
        res, to = ws_band_local(fd, lab, r0, r1, rank > 0, rank + 1 < world)
        a_res, a_to2 = comm.all_gather(torch.from_numpy(res)), comm.all_gather(torch.from_numpy(to))
        arr = bands.watershed_chain(a_res, a_to2, cols).numpy()
        side = 0 if rank > 0 else 1
        r = r0 if rank > 0 else r1 - 1
        for c in range(cols):
            v = arr[(rank * 2 + side) * cols + c]
            while v < 0:
                v = arr[-(v + 1)]
            assert v == ws[r, c]
        open(os.path.join(result_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_exchange_layer_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    world, rows, cols = 2, 256, 96
    mp.spawn(_gloo_worker, args=(world, _free_port(), rows, cols, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(world))
