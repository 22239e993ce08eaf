"""The certificates of tools/big_check.py against the oracle on a small raster: zero violations on the
oracle's outputs, and every certificate fires when its raster / table is corrupted.  (CPU tensors: the certificates
are plain torch arithmetic and never call the library.)"""
import importlib.util
import os

import numpy as np
import pytest
import torch

from malstroem_b200 import synth
from oracle import port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("big_check", os.path.join(ROOT, "tools", "big_check.py"))
bc = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bc)


class FakePipe(object):
    def __init__(self, dem):
        self.rows, self.cols = dem.shape
        filled = port.fill_terrain(dem)
        self.short, self.diag = port.minimum_safe_short_and_diag(dem)
        fnf = port.fill_terrain_no_flats(dem, self.short, self.diag)
        fd = port.terrain_flowdirection(fnf)
        acc = port.accumulated_flow(fd, fast=True)
        depths = filled - dem
        lab, n = port.connected_components(depths)
        ws = lab.copy()
        port.watersheds_from_labels(fd, ws, 0)
        st = port.label_stats(depths, lab, n)
        mi = port.label_min_index(fnf, lab, n)
        ma = port.label_max_index(acc, lab, n)
        cnt = np.zeros(n + 1, np.int64)
        c = port.label_count(ws)
        cnt[:c.size] = c
        self.nlabels = n
        self.dem = torch.from_numpy(dem)
        self.out = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in dict(
            filled=filled, depths=depths, fnf=fnf, flowdir=fd, accum=acc, labels=lab, wsheds=ws).items()}
        self.tables = {"st_min": st["min"], "st_max": st["max"], "st_sum": st["sum"], "st_count": st["count"].astype(np.int64),
                       "ws_count": cnt, "ppmin_value": mi["value"], "ppmin_row": mi["row"].astype(np.int64),
                       "ppmin_col": mi["col"].astype(np.int64), "ppmax_value": ma["value"],
                       "ppmax_row": ma["row"].astype(np.int64), "ppmax_col": ma["col"].astype(np.int64)}
        self.tables = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in self.tables.items()}

    def table(self, k):
        return self.tables[k][: self.nlabels + 1]


def flat(bad):
    out = {k: v for k, v in bad.items() if k != "tables"}
    out.update({"t_" + k: v for k, v in bad["tables"].items()})
    return out


@pytest.fixture(scope="module")
def pipe():
    return FakePipe(synth.fractal_dem(96, 130, seed=5))


def test_certificates_pass_on_oracle(pipe):
    assert pipe.nlabels > 5
    assert bc.certify(pipe, CH=40) == (0, 0, 0, float(pipe.rows * pipe.cols))
    bad = flat(bc.certify_more(pipe, pipe.short, pipe.diag, CH=40))
    assert all(v == 0 for v in bad.values()), bad


@pytest.mark.parametrize("what,key", [("fnf", "noflats"), ("flowdir", "d8"), ("labels", "cc_adjacent"),
                                      ("labels_bg", "cc_foreground"), ("swap", "cc_order"), ("st_max", "t_st_max"),
                                      ("ppmin_col", "t_ppmin"), ("ws_count", "t_ws_count"), ("st_sum", "t_st_sum")])
def test_certificates_fire(pipe, what, key):
    import copy
    p = copy.copy(pipe)
    p.out = {k: v.clone() for k, v in pipe.out.items()}
    p.tables = {k: v.clone() for k, v in pipe.tables.items()}
    lab = p.out["labels"]
    wet = torch.nonzero(lab > 0)
    r, c = [int(v) for v in wet[len(wet) // 2]]
    if what == "fnf":
        p.out["fnf"][40, 50] = torch.nextafter(p.out["fnf"][40, 50], torch.tensor(float("inf"), dtype=torch.float64))
    elif what == "flowdir":
        p.out["flowdir"][33, 44] = (int(p.out["flowdir"][33, 44]) + 1) % 8
    elif what == "labels":
        # split one cell of a multi-cell bluespot off into another label
        big = int(torch.bincount(lab.reshape(-1))[1:].argmax()) + 1
        rr, cc = [int(v) for v in torch.nonzero(lab == big)[0]]
        lab[rr, cc] = big % p.nlabels + 1
    elif what == "labels_bg":
        lab[r, c] = 0
    elif what == "swap":
        a, b = lab == 1, lab == 2
        lab[a], lab[b] = 2, 1
    elif what == "st_sum":
        p.tables[what][3] *= 1.0 + 1e-4
    else:
        p.tables[what][3] += 1
    bad = flat(bc.certify_more(p, p.short, p.diag, CH=40))
    assert bad[key] > 0, bad


def test_checksums_additive_over_bands(pipe):
    whole = bc.raster_checksums(pipe)

    class Band(object):
        pass
    tot = {k: torch.zeros((), dtype=torch.int64) for k in whole}
    for r0, r1 in ((0, 32), (32, 96)):
        b = Band()
        b.rows, b.cols = r1 - r0, pipe.cols
        b.out = {k: v[r0:r1] for k, v in pipe.out.items()}
        for k, v in bc.raster_checksums(b, row_offset=r0).items():
            tot[k] += v
    assert all(int(tot[k]) == int(whole[k]) for k in whole)
    pipe2 = Band()
    pipe2.rows, pipe2.cols = pipe.rows, pipe.cols
    pipe2.out = {k: v.clone() for k, v in pipe.out.items()}
    pipe2.out["accum"][17, 3] += 1
    assert int(bc.raster_checksums(pipe2)["accum"]) != int(whole["accum"])
