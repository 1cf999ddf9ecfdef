"""GPU: BASELINE config 2 -- an ensemble of 20x20 um networks swept across the percolation
threshold (measure_perc.py density sweep, slurm/batchmeasure.sh:6-14) in ONE ragged batch, and
config 1 (example.py: 5x5 um, n=300); a sample of the networks is re-done by the oracle."""
import numpy as np
import pytest

from oracle import cntnet_oracle as O

pytestmark = pytest.mark.gpu


def test_density_sweep_20um(engine):
    ns = [int(400 * d) for d in np.arange(1.0, 20.01, 0.25)]        # density 1 .. 20 /um^2, step 0.25: 77 networks
    seeds = np.arange(1, len(ns) + 1, dtype=np.uint64) * 7919
    res = engine.measure_ensemble(ns, 20.0, seeds, onoffmap=1)
    perc = res["percolating"]
    dens = np.array(ns) / 400.0
    assert not perc[dens <= 6].any() and perc[dens >= 13].all()      # threshold ~9 sticks/um^2
    assert np.all(res["status"] == 0)
    assert np.all(res["ion"][~perc] == 0) and np.all(res["ion"][perc] > 0)
    assert np.all(res["ioff_back"][perc] < res["ion"][perc])        # gating at +10 V lowers the current
    assert np.all(np.diff(res["njunctions"]) > -0.2 * res["njunctions"][1:])   # junctions grow ~ n^2
    # every 8th network again, through the oracle, from the very same sticks
    for k in range(0, len(ns), 8):
        b = engine.generate([ns[k]], 20.0, seeds=seeds[k:k + 1])
        s = engine.sticks_to_host(b)[0]
        m = O.measure_fullnet(s, onoffmap=1, refine=3)
        assert m["percolating"] == bool(perc[k])
        assert m["njunctions"] == res["njunctions"][k] and m["ntests"] == res["ntests"][k]
        assert m["nclust_ref"] == res["nclust"][k] and m["maxclust_ref"] == res["maxclust"][k]
        for key in ("ion", "ioff_back", "ioff_total", "ioff_partial"):
            if m["percolating"]:
                assert abs(res[key][k] - m[key]) <= 1e-9 * m[key], (k, key, res[key][k], m[key])
            else:
                assert res[key][k] == 0.0


def test_example_config_5um(engine):
    """example.py:2 -- n=300, 5x5 um; 64 independent seeds in one batch"""
    seeds = np.arange(100, 164, dtype=np.uint64)
    res = engine.measure_ensemble([300] * 64, 5.0, seeds, onoffmap=0)
    assert res["percolating"].mean() > 0.5
    for k in (0, 17, 63):
        s = engine.sticks_to_host(engine.generate([300], 5.0, seeds=seeds[k:k + 1]))[0]
        m = O.measure_fullnet(s, onoffmap=0, refine=3)
        assert m["percolating"] == bool(res["percolating"][k])
        for key in ("ion", "ioff_back", "ioff_total", "ioff_partial"):
            if m["percolating"]:
                assert abs(res[key][k] - m[key]) <= 1e-9 * m[key]


@pytest.mark.parametrize("n,onoffmap,element", [(120000, 1, "linexp"), (100000, 0, "fermidirac"), (90000, 2, "linexp")])
def test_full_size_variants(engine, n, onoffmap, element):
    """100x100 um at other densities / on-off maps / elements (slurm/batchmeasure_multi.sh sweeps density
    8..12): n = 120k does not fit the cluster kernel's on-chip capacity and takes the multi-launch
    solver automatically; all against the refined oracle from the same sticks"""
    from networksim_cntfet_b200 import elements
    el = elements.FermiDiracTransistor if element == "fermidirac" else elements.LinExpTransistor
    b = engine.generate([n], 100.0, seed=77)
    res = engine.measure_fullnet(b, onoffmap=onoffmap, element=el)
    rows = b.h_roff[1]
    from networksim_cntfet_b200 import _lib
    expect = "cluster" if rows <= _lib.load().cnt_pcg_cluster_max_rows() else "flat"
    assert engine.last_solver.startswith(expect), (engine.last_solver, rows)
    assert np.all(res["status"] == 0)
    s = engine.sticks_to_host(b)[0]
    m = O.measure_fullnet(s, onoffmap=onoffmap, element=element, refine=3)
    assert m["percolating"] == bool(res["percolating"][0])
    assert m["njunctions"] == res["njunctions"][0] and m["ntests"] == res["ntests"][0]
    for key in ("ion", "ioff_back", "ioff_total", "ioff_partial"):
        assert abs(res[key][0] - m[key]) <= 1e-9 * m[key], (key, res[key][0], m[key], res["iters"][:, 0])


def test_gate_configuration_sharding(engine):
    """one network spread over ranks by gate configuration (measure_perc.measure_fullnet_distributed):
    the shards of a 2-rank split add up to the unsharded result bit for bit"""
    full = engine.measure_fullnet(engine.generate([20000], 45.0, seeds=[11]), onoffmap=1)
    parts = [engine.measure_fullnet(engine.generate([20000], 45.0, seeds=[11]), onoffmap=1, slab_shard=(r, 2))
             for r in range(2)]
    assert full["percolating"][0]
    for k in ("ion", "ioff_back", "ioff_total", "ioff_partial"):
        assert parts[0][k][0] + parts[1][k][0] == full[k][0]
        assert (parts[0][k][0] == 0.0) != (parts[1][k][0] == 0.0)          # each configuration solved exactly once
    assert np.array_equal(parts[0]["iters"] + parts[1]["iters"], full["iters"])
    # without torch.distributed the distributed entry point degenerates to the plain one
    import networksim_cntfet_b200 as pkg
    from networksim_cntfet_b200 import measure_perc
    pkg.set_engine(engine)
    df, res = measure_perc.measure_fullnet_distributed(20000, 45.0, seed=11)
    assert res["ion"][0] == full["ion"][0] and df.ioff.tolist() == [full["ioff_back"][0], full["ioff_total"][0],
                                                                   full["ioff_partial"][0]]
