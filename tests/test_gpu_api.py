"""GPU: the drop-in host API (netsim / cnet / measure_perc mirrors) behaves like the reference:
same attributes, DataFrame columns, file formats, error behaviour -- values checked against the
golden vectors produced by the unmodified reference."""
import os

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture()
def api(engine, tmp_path, monkeypatch):
    import networksim_cntfet_b200 as pkg
    from networksim_cntfet_b200 import cnet, measure_perc, netsim
    pkg.set_engine(engine)
    monkeypatch.chdir(tmp_path)
    return netsim, cnet, measure_perc


def test_trivial_network_kat(api):
    """netsim.py:152-167 known-answer network through the class API"""
    netsim, cnet, _ = api
    g = load_golden("trivial")
    net = netsim.RandomCNTNetwork(n=2, scaling=1, seed=1, onoffmap=0)
    net.make_trivial_sticks()
    it = net.intersects
    assert list(it.columns) == ["stick1", "stick2", "x", "y", "kind"]
    assert it.stick1.tolist() == [0, 1, 2, 4] and it.stick2.tolist() == [2, 3, 3, 5]
    assert np.array_equal(it.x.values, g["jx"]) and np.array_equal(it.y.values, g["jy"])
    assert it.kind.tolist() == ["ms", "ms", "ss", "ss"]
    assert net.percolating
    assert list(net.sticks.columns) == ["xc", "yc", "angle", "length", "kind", "endarray", "cluster"]
    assert net.sticks.length.tolist() == [1.002, 1.001, 1, 1, 0.1, 0.1]
    gms, gss = np.exp(-5.0), np.exp(-1e-4)
    analytic = 0.1 / (2 / gms + 1 / gss)
    assert abs(sum(net.cnet.source_currents) - analytic) <= 1e-12 * analytic
    # gate sweep against the reference's own values (measure_perc.py:39-53 order)
    k = 0
    for gate in ("back", "partial", "total"):
        for vg in g["sweep_vg"]:
            cur = net.gate(vg, gate)
            assert abs(cur - g["sweep_current"][k]) <= 1e-9 * g["sweep_current"][k], (gate, vg)
            k += 1
    # networkx view with the reference's attributes (cnet.py:99-103,137-146)
    gr = net.cnet.graph
    assert sorted(gr.nodes()) == [0, 1, 2, 3]
    assert abs(gr.nodes[0]["voltage"] - 0.1) == 0 and gr.nodes[1]["voltage"] == 0
    e = gr.edges[2, 3]
    assert set(("conductance", "resistance", "current", "kind", "pos", "component")) <= set(e)
    net.label_clusters()
    assert net.sticks.cluster.tolist() == [0, 0, 0, 0, 1, 1] and net.clustersizes.tolist() == [4, 2]
    assert len(net.sticks.cluster.drop_duplicates()) == int(g["nclust_ref"])


@pytest.mark.parametrize("name,seed,n,om", [("s5_n300_seed1", 1, 300, 0), ("s5_n300_seed2_om1", 2, 300, 1),
                                              ("s5_n120_seed5_noperc", 5, 120, 0)])
def test_mt19937_replays_reference_seed(api, name, seed, n, om, capsys):
    """rng='mt19937' reproduces the sticks the reference draws for the same seed; electrode
    order is canonical (source index 0) where the reference's pandas sort may swap them"""
    netsim, _, _ = api
    g = load_golden(name)
    net = netsim.RandomCNTNetwork(n=n, scaling=5, seed=seed, onoffmap=om, rng="mt19937")
    st = net.sticks
    assert np.array_equal(st.length.values[2:], g["length"][2:])
    assert np.array_equal(st.xc.values[2:], g["xc"][2:]) and np.array_equal(st.angle.values[2:], g["angle"][2:])
    assert st.xc.values[0] == 0.01 and st.xc.values[1] == 0.99
    swap = g["xc"][0] == 0.99                                     # reference put the drain first
    ref1 = g["stick1"].copy()
    if swap:
        ref1 = np.where(ref1 <= 1, 1 - ref1, ref1)
    ref = sorted(zip(ref1.tolist(), g["stick2"].tolist()))
    got = sorted(zip(net.intersects.stick1.tolist(), net.intersects.stick2.tolist()))
    assert got == ref
    assert net.percolating == bool(g["percolating"])
    if net.percolating:
        assert abs(sum(net.cnet.source_currents) - float(g["isrc0"])) <= 1e-9 * float(g["isrc0"])
    else:
        assert "not conducting" in capsys.readouterr().out       # printed, not raised (netsim.py:215-218)
        assert not hasattr(net, "cnet")


def test_philox_class_and_save_load(api):
    netsim, _, _ = api
    net = netsim.RandomCNTNetwork(n=400, scaling=5, seed=12345, onoffmap=1)
    assert net.percolating and len(net.sticks) == 402
    assert net.sticks.xc[0] == 0.01 and net.sticks.kind[0] == "v" and np.all(np.diff(net.sticks.length.values[2:]) <= 0)
    ends = np.stack(net.sticks.endarray.values)
    assert ends.shape == (402, 2, 2)
    i0 = sum(net.cnet.source_currents)
    again = netsim.RandomCNTNetwork(n=400, scaling=5, seed=12345, onoffmap=1)
    assert sum(again.cnet.source_currents) == i0                    # same seed, same network, same bits
    os.makedirs("data", exist_ok=True)
    net.save_system()
    assert os.path.isfile(net.fname + "_sticks.csv") and os.path.isfile(net.fname + "_intersects.csv")
    loaded = netsim.RandomCNTNetwork(fname=os.path.basename(net.fname), directory="data", onoffmap=1)
    assert loaded.percolating and len(loaded.intersects) == len(net.intersects)
    # CSV round trip prints 17 significant digits: coordinates and currents survive exactly
    assert np.array_equal(loaded.intersects.x.values, net.intersects.x.values)
    assert abs(sum(loaded.cnet.source_currents) - i0) <= 1e-12 * i0
    net.label_clusters()                      # get_info needs clustersizes, as in the reference
    info = net.get_info()
    assert info[0] == 400 and info[7] is True


def test_load_reference_dump(api):
    """N2 parity input: `_sticks.csv` / `_intersects.csv` written by the REFERENCE's own save_system
    (netsim.py:228-236; tests/golden/make_golden.py dump_reference_system) load through load_system, and the GPU
    solve of that table reproduces the reference's currents, voltages and gate sweep; sheet conductance = I / Vds"""
    from conftest import GOLDEN_DIR
    netsim, _, _ = api
    ans = np.load(os.path.join(GOLDEN_DIR, "refdump_s5_n300_seed1_answers.npz"))
    net = netsim.RandomCNTNetwork(fname="refdump_s5_n300_seed1", directory=GOLDEN_DIR, onoffmap=int(ans["onoffmap"]))
    assert net.percolating == bool(ans["percolating"]) and len(net.sticks) == 302 and len(net.intersects) == 537
    # the reference under this pandas puts the x = 0.99 electrode at index 0: the dump is used as it is
    assert net.sticks.xc[0] == 0.99 and net.sticks.kind[0] == "v"
    i0 = float(sum(net.cnet.source_currents))
    assert abs(i0 - float(ans["isrc0"])) <= 1e-9 * float(ans["isrc0"])
    assert abs(net.cnet.sheet_conductance() - float(ans["isrc0"]) / 0.1) <= 1e-9 * float(ans["isrc0"]) / 0.1
    V = net.cnet.voltages()                    # table mode: the sticks of the percolating cluster, ascending index
    m = ~np.isnan(ans["V0"])
    assert len(V) == int(m.sum())
    assert np.max(np.abs(V - ans["V0"][m])) <= 1e-9 * 0.1
    cur = [net.gate(vg, gname) for gname in ("back", "partial", "total") for vg in (-10.0, 0.0, 10.0)]
    # the reference's own SuperLU answer at Vg = +10 is only ~2e-9 accurate
    assert np.all(np.abs(np.array(cur) - ans["sweep_current"]) <= 5e-9 * ans["sweep_current"])


def test_sheet_conductance_of_an_ensemble(api, engine):
    """sheet conductance G_sq = I_source / Vds for the four measure_fullnet gates against the oracle (1e-9) and
    as extra columns of the ensemble frame"""
    from oracle import cntnet_oracle as O
    _, _, mp = api
    seeds = np.arange(200, 206, dtype=np.uint64)
    res = engine.measure_ensemble([300] * 6, 5.0, seeds, onoffmap=1)
    assert res["sheet_conductance"].shape == (4, 6)
    for k in range(6):
        s = engine.sticks_to_host(engine.generate([300], 5.0, seeds=seeds[k:k + 1]))[0]
        m = O.measure_fullnet(s, onoffmap=1, refine=3)
        for q, key in enumerate(("ion", "ioff_back", "ioff_total", "ioff_partial")):
            want = m[key] / 0.1
            assert abs(res["sheet_conductance"][q, k] - want) <= 1e-9 * want if m["percolating"] else res["sheet_conductance"][q, k] == 0
    df = mp.measure_ensemble([300] * 6, 5, seeds, onoffmap=1, sheet=True)
    assert list(df.columns)[-2:] == ["gsheet_on", "gsheet_off"]
    assert np.allclose(df.gsheet_on.values, df.ion.values / 0.1, rtol=0, atol=0)


def test_measure_fullnet_and_async(api):
    _, _, mp = api
    df = mp.measure_fullnet(300, 5, seed=7, onoffmap=1)
    assert list(df.columns) == ["sticks", "size", "density", "nclust", "maxclust", "ion", "ioff", "gate", "fname",
                                "seed", "onoffmap", "runtime"]
    assert df.gate.tolist() in (["back", "total", "partial"], ["back"])
    # object route (save=True) gives the same numbers as the batched route
    os.makedirs("data", exist_ok=True)
    df2 = mp.measure_fullnet(300, 5, seed=7, onoffmap=1, save=True)
    assert df2.ion.tolist() == df.ion.tolist() and df2.ioff.tolist() == df.ioff.tolist()
    assert df2.nclust.tolist() == df.nclust.tolist() and df2.maxclust.tolist() == df.maxclust.tolist()
    assert os.path.isfile(df2.fname[0] + "_sticks.csv") and os.path.isfile(df2.fname[0] + "_data.csv")
    # density sweep across the percolation threshold (measure_perc.py:205-242), ragged sizes
    seeds = list(range(100, 112))
    out = mp.measure_async(4, 60, 30, 12, 5, onoffmap=[1], seeds=seeds)
    assert set(out.sticks) == {60 + 30 * i for i in range(12)}
    assert (out[out.ion == 0].gate == "back").all()
    perc = out[out.ion > 0]
    assert len(perc) % 3 == 0 and len(perc) > 0 and len(out[out.ion == 0]) > 0
    # deterministic in the seeds
    out2 = mp.measure_async(1, 60, 30, 12, 5, onoffmap=[1], seeds=seeds)
    assert out2.ion.tolist() == out.ion.tolist() and out2.ioff.tolist() == out.ioff.tolist()
    assert any(f.startswith("seeds_") for f in os.listdir("."))


def test_single_measure(api):
    _, cnet, mp = api
    data, fname = mp.single_measure(350, 5, seed=99, onoffmap=0, element=cnet.LinExpTransistor, vgnum=3,
                                    savedir="sweep")
    assert list(data.columns) == ["sticks", "scaling", "density", "current", "gatevoltage", "gate", "nclust",
                                  "maxclust", "fname", "onoffmap", "seed", "runtime", "element"]
    assert os.path.isfile(fname + "_data.csv")
    if len(data) == 9:
        assert data.gate.tolist() == ["back"] * 3 + ["partial"] * 3 + ["total"] * 3
        assert data.gatevoltage.tolist() == [-10.0, 0.0, 10.0] * 3
        back = data.current.values[:3]
        assert back[0] > back[1] > back[2] > 0                      # conductance falls with Vg (cnet.py:33-38)
        assert data.current.values[1] == data.current.values[4] == data.current.values[7]   # Vg=0 in every gate


def test_conduction_network_from_networkx(api):
    """reference signature ConductionNetwork(graph, ground_nodes, voltage_sources): a resistor
    ladder built from element objects (cnet.py:89-97)"""
    import networkx as nx
    _, cnet, _ = api
    G = nx.Graph()
    # nodes 0 (source) - 2 - 3 - 1 (ground) in series, plus a parallel branch 2 - 4 - 3
    for (u, v, R) in [(0, 2, 1.0), (2, 3, 2.0), (3, 1, 1.0), (2, 4, 1.0), (4, 3, 1.0)]:
        G.add_edge(u, v, component=cnet.Resistor(R), kind="ss", pos=[0.5, 0.5])
    net = cnet.ConductionNetwork(G, [1], [[0, 0.1]])
    net.update()
    req = 1.0 + 1.0 + 1.0        # 2||(1+1) = 1
    assert abs(sum(net.source_currents) - 0.1 / req) <= 1e-12
    V = net.voltages()            # canonical order: nodes 0,1,2,3,4
    assert np.allclose(V, [0.1, 0.0, 0.1 - 0.1 / 3, 0.1 / 3, 0.05], atol=1e-13)
    assert abs(net.graph.edges[2, 3]["current"] - (0.1 / 3) / 2.0) <= 1e-13
    x = net.solve_mna()
    assert len(x) == 5 and abs(x[-1] - 0.1 / req) <= 1e-12
    A = net.make_A(net.make_G())
    assert A.shape == (5, 5)


def test_gate_sweep_batched_equals_sequential(api):
    """config 4 (slurm/gatesweep.sh: --vgnum=11): the batched sweep gives the same currents as
    gate() step by step, and the reference's own sweep values for a replayed seed"""
    netsim, _, mp = api
    g = load_golden("s5_n300_seed1")
    net = netsim.RandomCNTNetwork(n=300, scaling=5, seed=1, onoffmap=0, rng="mt19937")
    cur = net.gate_sweep(g["sweep_vg"])
    ref = g["sweep_current"]
    assert len(cur) == 9
    for k in range(9):
        tol = 5e-9 if (k < 3 and g["sweep_vg"][k % 3] > 0) else 1e-9      # back gate at +10 V: reference's own error
        assert abs(cur[k] - ref[k]) <= tol * ref[k], (k, cur[k], ref[k])
    assert net.gatetype == "total" and net.gatevoltage == g["sweep_vg"][-1]
    assert abs(sum(net.cnet.source_currents) - cur[-1]) == 0
    seq = [net.gate(vg, gate) for gate in ("back", "partial", "total") for vg in g["sweep_vg"]]
    assert np.allclose(seq, cur, rtol=1e-12, atol=0)
    # 11-point sweep on a Philox device through the measure_perc driver
    data, fname = mp.single_measure(400, 5, seed=4242, onoffmap=0, vgnum=11, savedir="sweep11")
    if len(data) == 33:
        assert data.gate.tolist() == ["back"] * 11 + ["partial"] * 11 + ["total"] * 11
        back = data.current.values[:11]
        assert np.all(np.diff(back) < 0) and np.all(back > 0)
