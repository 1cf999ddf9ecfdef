"""CPU: the two exact reductions applied before assembly (leaf pruning, series elimination of an
independent set of degree-2 sticks -- csrc/prune.cu) restated in numpy (tests/reductions_ref.py)
and validated against the goldens of the unmodified reference: solving the REDUCED system and
restoring the removed sticks reproduces the reference's voltages."""
import numpy as np
import pytest
import scipy.sparse as sp
from scipy.sparse.linalg import spsolve

from conftest import FULLNET_GATES, golden_element, golden_names, load_golden
from oracle import cntnet_oracle as O
from reductions_ref import series_set, two_core

PERC = [n for n in golden_names() if "noperc" not in n]
VDS = 0.1


@pytest.mark.parametrize("name", PERC)
def test_reduced_system_reproduces_reference_voltages(name):
    g = load_golden(name)
    n = len(g["xc"])
    s1, s2 = g["stick1"].astype(np.int64), g["stick2"].astype(np.int64)
    if not (g["label"][1] == g["label"][0]):
        pytest.skip("not percolating")
    comp = g["label"] == 0
    alive, attach = two_core(n, s1, s2, comp)
    elim = series_set(n, s1, s2, alive)
    # structure: eliminated sticks form an independent set of degree-2 sticks with surviving neighbours
    for s, (a, c, e1, e2) in elim.items():
        assert alive[s] and s >= 2 and a >= 2 and c >= 2 and a not in elim and c not in elim
        assert {int(s1[e1]), int(s2[e1])} == {s, a} and {int(s1[e2]), int(s2[e2])} == {s, c}
    keep = alive.copy()
    keep[list(elim)] = False
    keep[:2] = False
    rows = np.flatnonzero(keep)
    rid = np.full(n, -1)
    rid[rows] = np.arange(len(rows))
    for tag, gate, vg in FULLNET_GATES:
        cond = O.conductance(g["kind2"], O.gate_voltages(g["jx"], g["jy"], gate, vg), int(g["onoffmap"]),
                             golden_element(g))
        # junctions of the reduced network: real ones between surviving sticks + one virtual per eliminated stick
        m = alive[s1] & alive[s2]
        for s in elim:
            m &= (s1 != s) & (s2 != s)
        a, c, ge = list(s1[m]), list(s2[m]), list(cond[m])
        for s, (na, nc, e1, e2) in elim.items():
            a.append(na)
            c.append(nc)
            ge.append(cond[e1] * cond[e2] / (cond[e1] + cond[e2]))
        a, c, ge = np.array(a), np.array(c), np.array(ge)
        diag = np.zeros(n)
        np.add.at(diag, a, ge)
        np.add.at(diag, c, ge)
        inner = (a >= 2) & (c >= 2)
        A = sp.coo_matrix((np.r_[-ge[inner], -ge[inner]], (np.r_[rid[a[inner]], rid[c[inner]]],
                                                           np.r_[rid[c[inner]], rid[a[inner]]])),
                          shape=(len(rows), len(rows))).tocsr() + sp.diags(diag[rows])
        rhs = np.zeros(n)
        src = a == 0                                  # stick 0 is always the first stick of its junctions
        np.add.at(rhs, c[src], VDS * ge[src])
        x = spsolve(A.tocsc(), rhs[rows])
        V = np.full(n, np.nan)
        V[0], V[1] = VDS, 0.0
        V[rows] = x
        for s, (na, nc, e1, e2) in elim.items():
            V[s] = (cond[e1] * V[na] + cond[e2] * V[nc]) / (cond[e1] + cond[e2])
        for s in np.flatnonzero(attach >= 0):
            V[s] = V[attach[s]]
        Vref = g["V0" if tag == "0" else "V" + tag]
        ok = ~np.isnan(Vref)
        assert np.array_equal(np.isnan(V), ~ok)
        assert np.max(np.abs(V[ok] - Vref[ok])) <= 1e-9 * VDS, (tag, np.max(np.abs(V[ok] - Vref[ok])))
    assert len(rows) + len(elim) + int((attach >= 0).sum()) + 2 == int(comp.sum())
