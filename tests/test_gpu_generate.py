"""GPU: counter-based stick placement (netsim.py:105-134 law) and the full pipeline on
GPU-generated sticks against the oracle run on the SAME (downloaded) stick tables."""
import numpy as np
import pytest

from oracle import cntnet_oracle as O

pytestmark = pytest.mark.gpu


def test_law_and_order(engine):
    n, scaling = 200000, 100.0
    b = engine.generate([n], scaling, seed=1234)
    s = engine.sticks_to_host(b)[0]
    assert len(s["xc"]) == n + 2
    # electrodes first: source (x=0.01) then drain (x=0.99), netsim.py:121-124
    assert s["xc"][0] == 0.01 and s["xc"][1] == 0.99 and s["yc"][0] == 0.5
    assert s["length"][0] == 100 and s["length"][1] == 100 and s["kind"][0] == 0 and s["kind"][1] == 0
    assert s["angle"][0] == np.pi / 2 - 1e-6
    assert s["orig"][0] == 0 and s["orig"][1] == n + 1
    body = slice(2, None)
    assert np.all(np.diff(s["length"][body]) <= 0)                 # sorted by length descending
    assert sorted(s["orig"].tolist()) == list(range(n + 2))        # a permutation of the pre-sort index
    # stable: equal lengths keep generation order
    eq = np.flatnonzero(np.diff(s["length"][body]) == 0)
    assert np.all(s["orig"][body][eq] < s["orig"][body][eq + 1])
    # law: |N(0.66,0.44)|/scaling, U(0,1)^2 centres, U(0,2pi) angle, P(metallic)=0.135
    L = s["length"][body] * scaling
    z = np.random.default_rng(0).normal(0.66, 0.44, 2_000_000)
    assert abs(L.mean() - np.abs(z).mean()) < 4e-3 and abs(L.std() - np.abs(z).std()) < 4e-3
    for k in ("xc", "yc"):
        v = s[k][body]
        assert 0 <= v.min() and v.max() < 1 and abs(v.mean() - 0.5) < 3e-3 and abs(v.var() - 1 / 12) < 2e-3
    a = s["angle"][body]
    assert 0 <= a.min() and a.max() < 2 * np.pi and abs(a.mean() - np.pi) < 2e-2
    pm = (s["kind"][body] == 1).mean()
    assert abs(pm - 0.135) < 4e-3 and set(np.unique(s["kind"][body])) == {1, 2}
    # independence of the draws (no stream overlap between the fields of a stick)
    o = np.argsort(s["orig"][body])
    for u, v in (("xc", "yc"), ("xc", "angle"), ("yc", "length"), ("angle", "length")):
        assert abs(np.corrcoef(s[u][body][o], s[v][body][o])[0, 1]) < 1e-2
    # end points follow netsim.py:97-103 (device sin/cos vs numpy: a few ulp)
    xa, ya, xb, yb = O.get_ends(s["xc"], s["yc"], s["angle"], s["length"])
    for got, want in ((s["xa"], xa), (s["ya"], ya), (s["xb"], xb), (s["yb"], yb)):
        assert np.max(np.abs(got - want)) < 1e-13


def test_counter_based_shard_independence(engine):
    """network k is the same whatever batch / offset it is generated in (ensembles shard
    over GPUs by network id; measure_perc.py:230-235 fans out the same way)"""
    ns = [300, 1000, 257, 5000]
    whole = engine.sticks_to_host(engine.generate(ns, 20.0, seed=77, net_id0=10))
    for k, n in enumerate(ns):
        one = engine.sticks_to_host(engine.generate([n], 20.0, seed=77, net_id0=10 + k))[0]
        for key in one:
            assert np.array_equal(one[key], whole[k][key]), (k, key)
    other = engine.sticks_to_host(engine.generate([300], 20.0, seed=78, net_id0=10))[0]
    assert not np.array_equal(other["xc"], whole[0]["xc"])


def test_constant_length_and_per_network_scaling(engine):
    b = engine.generate([500, 500], [5.0, 10.0], seed=3, l=1.0)
    s = engine.sticks_to_host(b)
    assert np.all(s[0]["length"][2:] == 1.0 / 5.0) and np.all(s[1]["length"][2:] == 1.0 / 10.0)
    # all body lengths tie: stable sort keeps generation order
    assert np.array_equal(s[0]["orig"][2:], np.arange(1, 501))


@pytest.mark.parametrize("ns,scaling", [([250, 300, 120, 400], 5.0), ([6400, 3600, 400], 20.0), ([100000], 100.0)])
def test_generated_pipeline_vs_oracle(engine, ns, scaling):
    b = engine.generate(ns, scaling, seed=2024)
    res = engine.measure_fullnet(b, onoffmap=1)
    tabs = engine.sticks_to_host(b)
    joff = b.joff.cpu().numpy()
    st1, st2 = b.stick1.cpu().numpy(), b.stick2.cpu().numpy()
    jx, jy, k2 = b.jx.cpu().numpy(), b.jy.cpu().numpy(), b.kind2.cpu().numpy()
    label = b.label.cpu().numpy()
    for i, s in enumerate(tabs):
        j = O.find_junctions(s)
        sl = slice(joff[i], joff[i + 1])
        assert np.array_equal(st1[sl], j["stick1"]) and np.array_equal(st2[sl], j["stick2"])
        assert np.array_equal(jx[sl], j["x"]) and np.array_equal(jy[sl], j["y"]) and np.array_equal(k2[sl], j["kind2"])
        assert int(res["ntests"][i]) == j["ntests"]
        c = O.label_clusters(len(s["xc"]), j["stick1"], j["stick2"], s["orig"])
        assert np.array_equal(label[b.h_soff[i]:b.h_soff[i + 1]], c["label"])
        assert res["nclust"][i] == c["nclust_ref"] and res["maxclust"][i] == c["maxclust_ref"]
        m = O.measure_fullnet(s, onoffmap=1, junctions=j, refine=3)
        assert bool(res["percolating"][i]) == m["percolating"]
        for k in ("ion", "ioff_back", "ioff_total", "ioff_partial"):
            if m["percolating"]:
                assert abs(res[k][i] - m[k]) <= 1e-9 * m[k], (i, k, res[k][i], m[k], res["iters"][:, i])
            else:
                assert res[k][i] == 0.0
