"""Parity of the §8(f) rows on the GPU (through the C ABI): next_downstream_label / pourpoint_network
(malstroem/algorithms/net.py:142-192) and Network.rain_event (malstroem/network.py:75-129) against outputs of the
reference itself (tests/golden/net188.npz, net_small.npz), the known answers of the reference's
tests/test_raster_net.py:8-21, and the CPU oracle on seeded inputs.  Labels and paths bit-exact; rain values
bit-exact under the same sum() flavour (north_star asks 1e-6 relative for volumes)."""
import numpy as np
import pytest

import net_cases
from conftest import _load
from malstroem_b200 import network, synth
from malstroem_b200.algorithms import net
from oracle import port

pytestmark = pytest.mark.gpu


def test_next_downstream_label_and_pourpoint_network_golden(dtm188):
    net_cases.check_net188(net, _load("net188.npz"), dtm188)


def test_geometric_pourpoint_network_golden(dtm188):
    net_cases.check_geometric_network(net, _load("net188.npz"), dtm188)


def test_net_small_cases(small_cases):
    net_cases.check_net_small(net, _load("net_small.npz"), small_cases)


def test_rain_events_golden():
    z, zs = _load("net188.npz"), _load("net_small.npz")
    net_cases.check_rain(network.rain_events_arrays, z, zs, exact=True)


def test_network_class_matches_reference_values():
    z = _load("net188.npz")
    nodes = [dict(nodeid=int(i), dstrnodeid=None if p == -1 else int(z["nodes_id"][p]), wshed_area=float(a),
                  bspot_vol=float(c))
             for i, p, a, c in zip(z["nodes_id"], z["nodes_parent"], z["nodes_area"], z["nodes_cap"])]
    nw = network.Network()
    nw.add_nodes(nodes)
    assert len(nw.root_nodes) == int((z["nodes_parent"] == -1).sum())
    for e, mm in enumerate(z["events"]):
        evs = nw.rain_event(mm)
        assert [d["nodeid"] for d in evs] == z["nodes_id"][z["nodes_order"]].tolist()      # the reference's order
        assert all(type(d["spillv"]) is (float if d["spillv"] else int) for d in evs)      # max(0, x): the int 0
        ev = {d["nodeid"]: d for d in evs}
        assert len(ev) == int(z["nodes_present"].sum())
        for k, nid in enumerate(z["nodes_id"]):
            d = ev[int(nid)]
            assert d["rainv"] == z["nodes_rainv"][e, k] and d["spillv"] == z["nodes_spillv"][e, k]
            assert d["v"] == z["nodes_v"][e, k]
            assert (d["pctv"] is None and np.isnan(z["nodes_pctv"][e, k])) or d["pctv"] == z["nodes_pctv"][e, k]


@pytest.mark.parametrize("size,seed", [(512, 1), (1024, 2)])
def test_network_on_fractal_vs_oracle(size, seed):
    """Forest walk (many pour points, background given) and plain walk against the oracle on a fractal DEM."""
    dem = synth.fractal_dem(size, size, seed=seed)
    filled = port.fill_terrain(dem)
    short, diag = port.minimum_safe_short_and_diag(dem)
    fnf = port.fill_terrain_no_flats(dem, short, diag)
    fd = port.terrain_flowdirection(fnf)
    lab, n = port.connected_components(filled - dem)
    mi = port.label_min_index(fnf, lab, n)
    cells = list(zip(mi["row"].tolist(), mi["col"].tolist()))
    for bg in (0, None):
        want = port.pourpoint_network(fd, lab, cells, bg)
        got = net.pourpoint_network(fd, lab, cells, bg)
        assert got == want
    got = net.pourpoint_network(fd, lab.astype(np.int64), cells[:50], 0)       # int64 labels, plain walk (< 64)
    assert got == want_first(port, fd, lab, cells[:50])
    # rain on the resulting node table
    parent = np.array([-1 if d["downstream_id"] is None else d["downstream_id"]
                       for d in port.pourpoint_network(fd, lab, cells, 0)])
    ws = lab.copy()
    port.watersheds_from_labels(fd, ws, 0)
    area = port.label_count(ws).astype(np.float64) * 0.16
    area = np.pad(area, (0, n + 1 - area.size))
    cap = port.label_stats(filled - dem, lab, n)["sum"] * 0.16
    mm = [10, 30, 100]
    want = port.rain_events(parent, area, cap, mm, 1)
    got = network.rain_events_arrays(parent, area, cap, mm, 1)
    for k in ("rainv", "spillv", "v", "pctv"):
        assert np.array_equal(got[k], want[k], equal_nan=True), k
    assert got["present"].all()


def want_first(port_, fd, lab, cells):
    return port_.pourpoint_network(fd, lab, cells, 0)


def test_pipeline_network_matches_functions():
    """ms_bluespot_network_dev on the device-resident tables == the function-level chain on the same rasters."""
    import torch
    from malstroem_b200.pipeline import RasterPipeline, synth_fractal
    size = 1024
    pipe = RasterPipeline(size, size)
    pipe.run(synth_fractal(size, size, seed=3))
    mm = [10.0, 30.0, 100.0]
    res = pipe.network(cell_area=0.16, events_mm=mm)
    torch.cuda.synchronize()
    n = pipe.nlabels
    fd = pipe.out["flowdir"].cpu().numpy()
    lab = pipe.out["labels"].cpu().numpy()
    cells = list(zip(pipe.table("ppmin_row").cpu().tolist(), pipe.table("ppmin_col").cpu().tolist()))
    want = port.pourpoint_network(fd, lab, cells, 0)
    parent = np.array([-1 if d["downstream_id"] is None else d["downstream_id"] for d in want])
    assert np.array_equal(res["parent"].cpu().numpy(), parent)
    area = pipe.table("ws_count").cpu().numpy().astype(np.float64) * 0.16
    cap = pipe.table("st_sum").cpu().numpy() * 0.16
    ref = port.rain_events(parent, area, cap, mm, network.SUM_MODE)
    for k in ("rainv", "spillv", "v", "pctv"):
        assert np.array_equal(res[k].cpu().numpy(), ref[k], equal_nan=True), k
    assert n > 1000


def test_rain_edge_cases():
    # empty network, single root, unknown downstream id, a two-node cycle next to a proper tree
    out = network.rain_events_arrays(np.zeros(0, np.int32), [], [], [10.0])
    assert out["rainv"].shape == (1, 0)
    out = network.rain_events_arrays([-1], [100.0], [0.0], [10.0, 20.0])
    assert out["present"].tolist() == [True] and np.isnan(out["pctv"]).all()
    assert out["spillv"][:, 0].tolist() == [1.0, 2.0] and out["v"][:, 0].tolist() == [0.0, 0.0]
    parent = np.array([-1, 0, -2, 2, 5, 4, 0])
    area = np.arange(1, 8) * 100.0
    cap = np.array([1.0, 0.5, 0.2, 0.0, 3.0, 1.0, 0.25])
    want = port.rain_events(parent, area, cap, [10.0], 1)
    got = network.rain_events_arrays(parent, area, cap, [10.0], 1)
    assert got["present"].tolist() == want["present"].tolist() == [True, True, False, False, False, False, True]
    m = want["present"]
    for k in ("rainv", "spillv", "v", "pctv"):
        assert np.array_equal(got[k][:, m], want[k][:, m], equal_nan=True)
    with pytest.raises(ValueError):
        network.rain_events_arrays([5], [1.0], [1.0], [10.0])          # parent index out of range


def test_net_error_behaviour():
    fd = np.zeros((5, 5), np.uint8)
    lab = np.zeros((5, 5), np.int32)
    with pytest.raises(ValueError):
        net.pourpoint_network(fd.astype(np.int32), lab, [(1, 1)], 0)
    with pytest.raises(IndexError):
        net.next_downstream_label(fd, lab, (7, 1), 0)
    assert net.next_downstream_label(fd, lab, (-1, 2), 0, geometry=True) == (None, [])
    assert net.pourpoint_network(fd, lab, [], 0) == []
    cyc = np.array([[2, 6]], np.uint8)                                  # two cells pointing at each other
    with pytest.raises(RuntimeError):
        net.next_downstream_label(cyc, np.zeros((1, 2), np.int32), (0, 0), 0)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_arbitrary_labels_and_nodir_cells_vs_oracle(seed):
    """Function-level semantics away from the tool chain: labels that are NOT connected components (speckle, stripes:
    paths re-enter the start cell's own label), no-direction cells in the interior, start cells anywhere, background
    given / None, int32 / int64 labels, few (plain walk) and many (forest walk) start cells, with geometry."""
    rng = np.random.default_rng(seed)
    rows, cols = 120 + 17 * seed, 150 - 11 * seed
    dem = synth.fractal_dem(rows, cols, seed=40 + seed)
    short, diag = port.minimum_safe_short_and_diag(dem)
    fd = port.terrain_flowdirection(port.fill_terrain_no_flats(dem, short, diag))
    fd[rng.random(fd.shape) < 0.02] = 8                                    # interior cells without a direction
    lab = np.where(rng.random(fd.shape) < 0.3, rng.integers(1, 6, fd.shape), 0).astype(np.int32)     # speckle
    lab[(np.arange(rows) // 7) % 3 == 0, :] = 7                            # stripes of one label
    cells = [(int(r), int(c)) for r, c in zip(rng.integers(0, rows, 300), rng.integers(0, cols, 300))]
    for labels in (lab, lab.astype(np.int64)):
        for bg in (0, None, 7):
            for sub in (cells, cells[:40]):
                assert net.pourpoint_network(fd, labels, sub, bg) == port.pourpoint_network(fd, labels, sub, bg)
    for cell in cells[:25]:
        for bg in (0, None):
            got = net.next_downstream_label(fd, lab, cell, bg, geometry=True)
            want = port.next_downstream_label(fd, lab, cell, bg, geometry=True)
            assert got[0] == want[0] and [tuple(map(int, c)) for c in got[1]] == want[1]
