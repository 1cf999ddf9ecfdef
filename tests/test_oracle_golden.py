"""CPU: pin the numpy restatement (oracle/cntnet_oracle.py) against the golden vectors that
the UNMODIFIED reference produced (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from conftest import FULLNET_GATES, golden_element, golden_names, golden_sticks, load_golden
from oracle import cntnet_oracle as O

NAMES = golden_names()


def test_goldens_present():
    assert len(NAMES) >= 10


@pytest.mark.parametrize("name", NAMES)
def test_ends_restatement(name):
    g = load_golden(name)
    xa, ya, xb, yb = O.get_ends(g["xc"], g["yc"], g["angle"], g["length"])
    for a, b in zip((xa, ya, xb, yb), (g["xa"], g["ya"], g["xb"], g["yb"])):
        assert np.array_equal(a, b)          # netsim.py:97-103, bit-exact


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("pruner", ["kdtree", "brute"])
def test_junctions_bit_exact(name, pruner):
    g = load_golden(name)
    if pruner == "brute" and len(g["xc"]) > 4000:
        pytest.skip("all-pairs too large")
    j = O.find_junctions(golden_sticks(g), pruner=pruner)
    assert np.array_equal(j["stick1"], g["stick1"])
    assert np.array_equal(j["stick2"], g["stick2"])
    assert np.array_equal(j["x"], g["jx"])           # bit-exact coordinates
    assert np.array_equal(j["y"], g["jy"])
    assert np.array_equal(j["kind2"], g["kind2"])


@pytest.mark.parametrize("name", NAMES)
def test_clusters(name):
    g = load_golden(name)
    c = O.label_clusters(len(g["xc"]), g["stick1"], g["stick2"], g["orig"])
    assert np.array_equal(c["label"], g["label"])
    assert c["percolating"] == bool(g["percolating"])
    assert c["nclust_ref"] == int(g["nclust_ref"])       # measure_perc.py:76 quirk
    assert c["maxclust_ref"] == int(g["maxclust_ref"])   # measure_perc.py:78 quirk
    sizes = np.bincount(c["label"])
    assert sorted(sizes[sizes > 1].tolist()) == sorted(g["clustersizes"].tolist())


@pytest.mark.parametrize("name", [n for n in NAMES if "noperc" not in n and "n3600" not in n])
def test_solve_matches_reference(name):
    g = load_golden(name)
    n = len(g["xc"])
    el, om = golden_element(g), int(g["onoffmap"])
    for tag, gate, vg in FULLNET_GATES:
        vgj = O.gate_voltages(g["jx"], g["jy"], gate, vg)
        cond = O.conductance(g["kind2"], vgj, om, el)
        Gref = g["G0" if tag == "0" else "G" + tag]
        m = ~np.isnan(Gref)
        assert np.array_equal(cond[m], Gref[m])           # conductances bit-exact (cnet.py:33-41,60-66)
        r = O.solve_network(n, g["stick1"], g["stick2"], cond, g["label"], refine=3)
        Vref = g["V0" if tag == "0" else "V" + tag]
        Iref = g["I0" if tag == "0" else "I" + tag]
        isr = float(g["isrc0" if tag == "0" else "isrc" + tag])
        assert np.array_equal(np.isnan(r["V"]), np.isnan(Vref))
        mv = ~np.isnan(Vref)
        # the reference's own SuperLU solve of the indefinite MNA matrix carries up to ~4e-11
        # (norm-wise, relative to Vds) / 2e-9 (source current at Vg=+10) error against the
        # refined solution; see DESIGN.md "PCG accuracy"
        assert np.max(np.abs(r["V"][mv] - Vref[mv])) <= 1e-9 * 0.1
        assert np.all(np.abs(r["I"][m] - Iref[m]) <= 1e-9 * np.abs(Iref[m]) + cond[m] * 2e-10 * 0.1)
        tol = 1e-9 if gate != "back" else 5e-9
        assert abs(r["I_source"] - isr) <= tol * abs(isr)


def test_trivial_analytic():
    """netsim.py:152-167: two ms contacts in series with one ss junction."""
    g = load_golden("trivial")
    assert g["stick1"].tolist() == [0, 1, 2, 4] and g["stick2"].tolist() == [2, 3, 3, 5]
    assert g["label"].tolist() == [0, 0, 0, 0, 4, 4]
    gms, gss = np.exp(-0.5 * 0) * np.exp(-5.0), np.exp(-1e-5 * 0) * np.exp(-1e-4)
    analytic = 0.1 / (2 / gms + 1 / gss)
    cond = O.conductance(g["kind2"], np.zeros(4), 0, "linexp")
    r = O.solve_network(6, g["stick1"], g["stick2"], cond, g["label"])
    assert abs(r["I_source"] - analytic) <= 1e-15 * analytic * 10
    assert abs(float(g["isrc0"]) - analytic) <= 1e-14 * analytic


def test_sweep_restatement():
    """RandomCNTNetwork.gate sweep (netsim.py:266-277, measure_perc.py:39-53)."""
    g = load_golden("s5_n300_seed1")
    n = len(g["xc"])
    k = 0
    for gate in ("back", "partial", "total"):
        for vg in g["sweep_vg"]:
            vgj = O.gate_voltages(g["jx"], g["jy"], gate, vg)
            cond = O.conductance(g["kind2"], vgj, 0, "linexp")
            r = O.solve_network(n, g["stick1"], g["stick2"], cond, g["label"], refine=3)
            ref = g["sweep_current"][k]
            assert abs(r["I_source"] - ref) <= (5e-9 if (gate == "back" and vg > 0) else 1e-9) * abs(ref)
            k += 1


def test_generator_law():
    """make_sticks restatement reproduces the reference's MT19937 stream (netsim.py:105-125):
    the sorted lengths/centres of seed 1 equal the golden table's."""
    g = load_golden("s5_n300_seed1")
    t = O.sticks_with_ends(O.make_sticks_mt19937(300, "exp", 0.135, 5, 1))
    assert np.array_equal(np.sort(t["length"]), np.sort(g["length"]))
    assert np.array_equal(np.sort(t["xc"]), np.sort(g["xc"]))
    o = g["orig"]
    for k in ("xc", "yc", "angle", "length", "xa", "yb"):
        assert np.array_equal(t[k][o], g[k])
    s = O.canonical_sort(O.make_sticks_mt19937(300, "exp", 0.135, 5, 1))
    assert s["xc"][0] == 0.01 and s["xc"][1] == 0.99 and np.all(np.diff(s["length"][2:]) <= 0)


@pytest.mark.parametrize("name", [n for n in golden_names() if "noperc" not in n and "n3600" not in n and "n6400" not in n])
def test_exact_star_mesh_oracle_agrees_with_reference_goldens(name):
    """extended-precision star-mesh elimination of the ideal network (oracle.solve_network_exact) vs the
    voltages of the unmodified reference: on these small, well-conditioned networks both are right"""
    g = load_golden(name)
    for tag, gate, vg in FULLNET_GATES:
        cond = O.conductance(g["kind2"], O.gate_voltages(g["jx"], g["jy"], gate, vg), int(g["onoffmap"]),
                             golden_element(g))
        V = O.solve_network_exact(len(g["xc"]), g["stick1"], g["stick2"], cond, g["label"])
        Vref = g["V0" if tag == "0" else "V" + tag]
        ok = ~np.isnan(Vref)
        assert np.array_equal(np.isnan(V), ~ok)
        assert np.max(np.abs(V[ok] - Vref[ok])) <= 1e-9 * 0.1, (tag, np.max(np.abs(V[ok] - Vref[ok])))


def test_exact_oracle_on_a_floating_cluster():
    """50 sticks in a chain hanging on the electrodes by 1e-12 contacts: the exact oracle gives the
    analytic answer, the LU of the assembled matrix does not"""
    n = 52
    s1 = np.r_[0, np.arange(2, n - 1), 1]
    s2 = np.r_[2, np.arange(3, n), n - 1]
    order = np.lexsort((s2, s1))
    s1, s2 = s1[order], s2[order]
    g = np.where(s1 < 2, np.where(s1 == 0, 1e-12, 3e-12), 1.0)
    label = np.zeros(n, dtype=np.int64)
    V = O.solve_network_exact(n, s1, s2, g, label)
    rs, rd = 1 / 1e-12, 1 / 3e-12
    want = 0.1 * (1 - (rs + np.arange(n - 2)) / (rs + (n - 3) + rd))
    assert np.max(np.abs(V[2:] - want) / want) <= 1e-15
    lu = O.solve_network(n, s1, s2, g, label, refine=0)["V"]
    assert np.max(np.abs(lu[2:] - want) / want) > 1e-9
