"""GPU: the star-mesh direct solver (Engine.solver = "direct"; starmesh.py + csrc/starmesh.cu)
against the torch restatement of its numeric phase, the goldens of the unmodified reference, the
oracle and the iterative cluster solver."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import slab_ref  # noqa: E402
import starmesh_ref  # noqa: E402
from conftest import FULLNET_GATES, golden_element, golden_names, golden_sticks, load_golden  # noqa: E402
from networksim_cntfet_b200 import starmesh  # noqa: E402
from oracle import cntnet_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu
VDS = 0.1
PERC = [n for n in golden_names() if "noperc" not in n and "n3600" not in n]


@pytest.fixture()
def direct(engine):
    old = engine.solver
    engine.solver = "direct"
    yield engine
    engine.solver = old
    engine.direct_leak = True


def _element(name):
    from networksim_cntfet_b200 import elements
    return elements.FermiDiracTransistor if name == "fermidirac" else elements.LinExpTransistor


def test_cuda_numeric_equals_torch_restatement(engine):
    system = slab_ref.random_network_system(n=30, seed=4, nconf=3)
    rowptr, col, val, diag, rhs, mats = system
    R = diag.shape[1]
    rows = np.repeat(np.arange(R), np.diff(rowptr))
    c = diag + np.stack([np.bincount(rows, weights=v, minlength=R) for v in val])
    t = lambda a, dt, dev: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    plan_c = starmesh.symbolic(t(rowptr, torch.int64, "cpu"), t(col, torch.int64, "cpu"), R)
    want = starmesh_ref.solve(plan_c, t(val, torch.float64, "cpu"), t(c, torch.float64, "cpu"), t(rhs, torch.float64, "cpu"))
    dev = engine.device
    with engine._work():
        plan = starmesh.finalize(starmesh.symbolic(t(rowptr, torch.int64, dev), t(col, torch.int64, dev), R))
        gd = t(c, torch.float64, dev)              # all of the coupling through one "electrode"
        gs = torch.zeros_like(gd)
        x = starmesh.solve_cuda(plan, engine.lib, engine.stream_ptr, t(rowptr, torch.int32, dev), t(val, torch.float64, dev),
                                gs, gd, t(rhs, torch.float64, dev), leak=False)
    engine.sync()
    assert plan.nrounds == len(plan_c.rounds)      # same plan on both devices
    assert float((x.cpu() - want).abs().max()) <= 1e-13 * VDS


@pytest.mark.parametrize("dense_max", [40, 1000])
def test_cuda_dense_tail_equals_torch_restatement(engine, dense_max):
    """batch of two copies of a system; sparse rounds down to dense_max unknowns per network, then k_sm_dense"""
    system = slab_ref.random_network_system(n=16, seed=9, nconf=3)
    rowptr, col, val, diag, rhs, mats = system
    R = diag.shape[1]
    rows = np.repeat(np.arange(R), np.diff(rowptr))
    c = diag + np.stack([np.bincount(rows, weights=v, minlength=R) for v in val])
    rowptr2, col2 = np.r_[rowptr, rowptr[-1] + rowptr[1:]], np.r_[col, col + R]
    t = lambda a, dt, dev: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    out = {}
    for dev in ("cpu", engine.device):
        kw = dict(local_ids=t(np.r_[np.arange(R), np.arange(R)], torch.int64, dev),
                  net_ids=t(np.repeat([0, 1], R), torch.int64, dev), nnet=2, dense_max=dense_max)
        if dev == "cpu":
            plan = starmesh.symbolic(t(rowptr2, torch.int64, dev), t(col2, torch.int64, dev), 2 * R, **kw)
            out[dev] = starmesh_ref.solve(plan, t(np.c_[val, val], torch.float64, dev), t(np.c_[c, c], torch.float64, dev),
                                          t(np.c_[rhs, rhs], torch.float64, dev))
        else:
            with engine._work():
                plan = starmesh.finalize(starmesh.symbolic(t(rowptr2, torch.int64, dev), t(col2, torch.int64, dev), 2 * R, **kw))
                gd = t(np.c_[c, c], torch.float64, dev)
                out["gpu"] = starmesh.solve_cuda(plan, engine.lib, engine.stream_ptr, t(rowptr2, torch.int32, dev),
                                                 t(np.c_[val, val], torch.float64, dev), torch.zeros_like(gd), gd,
                                                 t(np.c_[rhs, rhs], torch.float64, dev), leak=False)
            engine.sync()
            assert plan.dense is not None and plan.dense["nmax"] <= dense_max
    x = out["gpu"].cpu()
    assert float((x - out["cpu"]).abs().max()) <= 1e-13 * VDS
    assert torch.equal(x[:, :R], x[:, R:])


def test_dense_tail_switch_does_not_change_the_answer(direct):
    """sparse rounds to the end vs the dense tail: same currents to rounding (20x20 um batch)"""
    res = {}
    for dm in (0, 384):
        direct.direct_dense_max = dm
        try:
            res[dm] = direct.measure_fullnet(direct.generate([6400, 5000, 3000], 20.0, seed=8), onoffmap=1)
            assert (direct.last_direct["dense_nmax"] > 0) == (dm > 0)
        finally:
            direct.direct_dense_max = 384
    for key in ("ion", "ioff_back", "ioff_total", "ioff_partial"):
        assert np.all(np.abs(res[0][key] - res[384][key]) <= 1e-12 * np.abs(res[0][key]))


@pytest.mark.parametrize("name", PERC)
def test_fullnet_direct_vs_goldens(direct, name):
    """measure_fullnet with the direct solver: V, |I|, I_source against the refined oracle and the raw reference"""
    g = load_golden(name)
    b = direct.batch_from_host([golden_sticks(g)])
    res = direct.measure_fullnet(b, onoffmap=int(g["onoffmap"]), element=_element(golden_element(g)), want_fields=True)
    assert direct.last_solver == "direct" and res["percolating"][0] and np.all(res["status"] == 0)
    f = res["fields"]
    for q, (tag, gate, vg) in enumerate(FULLNET_GATES):
        vgj = O.gate_voltages(g["jx"], g["jy"], gate, vg)
        cond = O.conductance(g["kind2"], vgj, int(g["onoffmap"]), golden_element(g))
        r = O.solve_network(len(g["xc"]), g["stick1"], g["stick2"], cond, g["label"], refine=3)
        V = f["V"][q].cpu().numpy()[:b.S]
        m = ~np.isnan(r["V"])
        assert np.array_equal(np.isnan(V), ~m)
        assert np.max(np.abs(V[m] - r["V"][m])) <= 1e-9 * VDS, (tag, np.max(np.abs(V[m] - r["V"][m])))
        isrc = res["isrc_all"][q, 0]
        assert abs(isrc[2] - r["I_source"]) <= 1e-9 * r["I_source"], (tag, isrc, r["I_source"])
        Vref = g["V0" if tag == "0" else "V" + tag]
        assert np.max(np.abs(V[m] - Vref[m])) <= 1e-9 * VDS


def test_direct_full_size_vs_cluster_and_oracle(direct):
    """3 networks of 100x100 um: currents against the iterative cluster solver and the refined oracle;
    voltages of the assembled FP64 matrix (leak on) against the cluster solver's"""
    direct.solver = "cluster"
    ref = direct.measure_fullnet(direct.generate([100000] * 3, 100.0, seed=31), onoffmap=1, want_fields=True)
    direct.solver = "direct"
    b = direct.generate([100000] * 3, 100.0, seed=31)
    res = direct.measure_fullnet(b, onoffmap=1, want_fields=True)
    assert direct.last_solver == "direct" and np.all(res["status"] == 0)
    info = direct.last_direct
    assert info["rounds"] < 200 and info["nnzL"] < 3 * b.NNZ
    for key in ("ion", "ioff_back", "ioff_total", "ioff_partial"):
        assert np.all(np.abs(res[key] - ref[key]) <= 1e-9 * ref[key]), key
    V, Vr = res["fields"]["V"].cpu().numpy(), ref["fields"]["V"].cpu().numpy()
    ok = ~np.isnan(Vr)
    assert np.array_equal(np.isnan(V), ~ok)
    err = np.abs(V - Vr)[ok].reshape(4, -1).max(1) / VDS
    assert np.all(err <= 1e-9), err        # same matrix, two solvers (PCG at rtol 1e-15 is ~1e-10 from its exact solution)
    m = O.measure_fullnet(direct.sticks_to_host(b)[0], onoffmap=1, refine=3)
    for key in ("ion", "ioff_back", "ioff_total", "ioff_partial"):
        assert abs(res[key][0] - m[key]) <= 1e-9 * m[key]


def test_direct_is_deterministic_and_plan_is_reused(direct):
    b = direct.generate([6400, 3000, 6400], 20.0, seed=8)
    r1 = direct.measure_fullnet(b, onoffmap=1, want_fields=True)
    plan = b._sm_plan
    r2 = direct.measure_fullnet(b, onoffmap=1, want_fields=True)
    assert b._sm_plan is plan
    assert torch.equal(r1["fields"]["x"], r2["fields"]["x"])
    b2 = direct.generate([6400, 3000, 6400], 20.0, seed=8)
    r3 = direct.measure_fullnet(b2, onoffmap=1, want_fields=True)
    assert torch.equal(r1["fields"]["x"], r3["fields"]["x"])
