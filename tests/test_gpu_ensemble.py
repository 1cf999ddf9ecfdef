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
    8..12) through the default (direct) solver; all against the refined oracle from the same sticks"""
    from networksim_cntfet_b200 import elements
    el = elements.FermiDiracTransistor if element == "fermidirac" else elements.LinExpTransistor
    b = engine.generate([n], 100.0, seed=77)
    res = engine.measure_fullnet(b, onoffmap=onoffmap, element=el)
    rows = b.h_roff[1]
    from networksim_cntfet_b200 import _lib
    assert engine.last_solver == "direct", (engine.last_solver, rows)      # sparse pattern: star-mesh direct solver
    assert np.all(res["status"] == 0)
    s = engine.sticks_to_host(b)[0]
    j = O.find_junctions(s)
    m = O.measure_fullnet(s, onoffmap=onoffmap, element=element, refine=3, junctions=j)
    assert m["percolating"] == bool(res["percolating"][0])
    assert m["njunctions"] == res["njunctions"][0] and m["ntests"] == res["ntests"][0]
    # the junction SET and coordinates, bit for bit, and the cluster labels (not only the counts)
    J = int(res["njunctions"][0])
    assert np.array_equal(b.stick1[:J].cpu().numpy(), j["stick1"]) and np.array_equal(b.stick2[:J].cpu().numpy(), j["stick2"])
    assert np.array_equal(b.jx[:J].cpu().numpy(), j["x"]) and np.array_equal(b.jy[:J].cpu().numpy(), j["y"])
    assert np.array_equal(b.kind2[:J].cpu().numpy(), j["kind2"])
    c = O.label_clusters(len(s["xc"]), j["stick1"], j["stick2"], s["orig"])
    assert np.array_equal(b.label[:len(s["xc"])].cpu().numpy(), c["label"])
    for key in ("ion", "ioff_back", "ioff_total", "ioff_partial"):
        assert abs(res[key][0] - m[key]) <= 1e-9 * m[key], (key, res[key][0], m[key], res["iters"][:, 0])


def test_full_size_properties(engine):
    """BASELINE size (3 networks of 100x100 um, n = 100k, measure_fullnet): size-independent
    properties of the answer, checked from the returned fields alone (torch ops on V, I, g and the
    junction table -- nothing of the solver is reused):
    Kirchhoff's current law at every stick of the component (dangling sticks included), the three
    source-current estimators agree, maximum principle 0 <= V <= Vds, pruned sticks sit at their
    attachment point, gating can only lower the current (Rayleigh monotonicity: every LinExp
    conductance falls with Vg), currents are unsigned, and the run is bit-reproducible."""
    import torch
    b = engine.generate([100000, 100000, 100000], 100.0, seed=2024)
    res = engine.measure_fullnet(b, onoffmap=1, want_fields=True)
    assert np.all(res["status"] == 0) and np.all(res["percolating"])
    f = res["fields"]
    S, J = b.S, b.J
    soff = torch.as_tensor(b.h_soff, device=engine.device, dtype=torch.int64)
    net_of_j = torch.bucketize(torch.arange(J, device=engine.device), b.joff[1:].to(torch.int64), right=True)
    a = b.stick1[:J].to(torch.int64) + soff[net_of_j]
    c = b.stick2[:J].to(torch.int64) + soff[net_of_j]
    local = torch.arange(S, device=engine.device) - soff[torch.bucketize(torch.arange(S, device=engine.device), soff[1:], right=True)]
    for q in range(4):
        V, I, g = f["V"][q][:S], f["I"][q][:J], f["g"][q][:J]
        inc = ~torch.isnan(V)
        assert torch.equal(inc, b.label[:S] == 0)                       # exactly the percolating component
        e = inc[a] & inc[c]
        assert torch.equal(~torch.isnan(I), e)
        Vz = torch.nan_to_num(V)
        flow = torch.where(e, g * (Vz[a] - Vz[c]), torch.zeros_like(g))  # signed current a -> c
        assert torch.all(I[e] >= 0) and torch.allclose(I[e], flow[e].abs(), rtol=0, atol=0)
        net = torch.zeros(S, dtype=torch.float64, device=engine.device)
        net.index_add_(0, a, flow)
        net.index_add_(0, c, -flow)
        inner = inc & (local >= 2)
        gsum = torch.zeros(S, dtype=torch.float64, device=engine.device)
        ge = torch.where(e, g, torch.zeros_like(g))
        gsum.index_add_(0, a, ge)
        gsum.index_add_(0, c, ge)
        # KCL: what flows into a stick flows out (residual relative to the current through the network)
        for k in range(b.B):
            m = inner & (torch.arange(S, device=engine.device) >= b.h_soff[k]) & (torch.arange(S, device=engine.device) < b.h_soff[k + 1])
            itot = res["isrc_all"][q, k, 2]
            # node-wise: the imbalance is what a voltage error of 1e-11 Vds at that stick would cause
            # (sum of its conductances x 1e-12 V); the parity bar on V itself is 1e-9 Vds
            assert bool(torch.all(net[m].abs() <= 1e-12 * gsum[m])), (q, k, float((net[m].abs() / gsum[m]).max()), itot)
            src, drn, pw = res["isrc_all"][q, k]
            assert abs(src - pw) <= 1e-6 * pw and abs(drn - pw) <= 1e-7 * pw
            # current leaving the source electrode == current entering the drain (from the fields)
            assert abs(float(net[b.h_soff[k]]) - pw) <= 1e-6 * pw and abs(float(-net[b.h_soff[k] + 1]) - pw) <= 1e-7 * pw
        # maximum principle up to the attainable accuracy of V (see the comparison with the oracle below)
        assert float(V[inc].min()) >= -1e-9 and float(V[inc].max()) <= 0.1 + 1e-9
        if b.attach is not None:      # exact reductions applied (iterative solvers): dangling sticks carry no current
            pr = b.attach[:S] >= 0
            assert int(pr.sum()) > 10000 and torch.equal(V[pr], V[b.attach[:S][pr].to(torch.int64)])
    # a local gate lowers the conductances of a subset of the junctions the back gate lowers
    for key in ("ioff_partial", "ioff_total"):
        assert np.all(res["ion"] > res[key]) and np.all(res[key] > res["ioff_back"])
    assert np.all(res["ioff_back"] > 0)
    # voltages of the first network against the refined oracle (SuperLU + 3 refinement steps, itself
    # within 3e-13 Vds of an extended-precision solve).  At Vg = 0 and under the local gates: 1e-9 Vds.
    # Under the back gate at +10 V the whole interior floats on ~100 contacts of e^-10: the rounding of
    # the 58k diagonal sums (1e-16 each) is a spurious leak comparable to that stiffness, so the
    # solution of the FP64 matrix is itself defined only to a few 1e-9 Vds -- two summation orders of
    # the diagonal differ by 3e-9, and the reference's own SuperLU answer is 5e-9 off.
    # Currents are insensitive (1e-12).
    s = engine.sticks_to_host(b)[0]
    jn = O.find_junctions(s)
    lab = O.label_clusters(len(s["xc"]), jn["stick1"], jn["stick2"])["label"]
    for q, (gate, vgv, tol) in enumerate(((None, 0.0, 1e-9), ("back", 10.0, 8e-9), ("total", 10.0, 1e-9), ("partial", 10.0, 1e-9))):
        cond = O.conductance(jn["kind2"], O.gate_voltages(jn["x"], jn["y"], gate, vgv), 1, "linexp")
        r = O.solve_network(len(s["xc"]), jn["stick1"], jn["stick2"], cond, lab, refine=3)
        V = f["V"][q][:b.h_soff[1]].cpu().numpy()
        m = ~np.isnan(r["V"])
        err = np.abs(V[m] - r["V"][m])
        assert err.max() <= tol * 0.1, (gate, err.max() / 0.1)
        assert abs(res["isrc_all"][q, 0, 2] - r["I_source"]) <= 1e-12 * r["I_source"]
    again = engine.measure_fullnet(engine.generate([100000, 100000, 100000], 100.0, seed=2024), onoffmap=1)
    for key in ("ion", "ioff_back", "ioff_total", "ioff_partial", "iters"):
        assert np.array_equal(res[key], again[key]), key


def test_pipelined_batches_equal_sequential(engine, monkeypatch):
    """measure_perc.PIPELINE: the batches of one call alternate between two engines on two host
    threads (Engine.run_pipelined); every network is seeded individually -> identical rows"""
    import networksim_cntfet_b200 as pkg
    from networksim_cntfet_b200 import measure_perc
    pkg.set_engine(engine)
    ns = [3000, 3600, 4000, 4400, 5200, 6400, 3800, 4100]
    seeds = list(range(501, 509))
    monkeypatch.setattr(measure_perc, "BATCH_STICKS", 9000)        # 2 networks per batch -> 4+ batches
    monkeypatch.setattr(measure_perc, "PIPELINE_MIN_STICKS", 0)    # (small calls stay on one engine by default)
    monkeypatch.setattr(measure_perc, "PIPELINE", False)
    a = measure_perc.measure_ensemble(ns, 20, seeds, onoffmap=1)
    monkeypatch.setattr(measure_perc, "PIPELINE", True)
    b = measure_perc.measure_ensemble(ns, 20, seeds, onoffmap=1)
    assert len(a) == len(b) and len(a) >= len(ns)
    for col in ("sticks", "nclust", "maxclust", "ion", "ioff", "gate", "seed"):
        assert a[col].tolist() == b[col].tolist(), col


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
