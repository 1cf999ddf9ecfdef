"""CPU: the numpy restatement of the star-mesh direct solver (tests/starmesh_ref.py -- the checker of
csrc/starmesh_sym.cu + csrc/starmesh.cu on the GPU, tests/test_gpu_direct.py) against scipy's direct
solve, and the invariants every plan must satisfy."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp
from scipy.sparse.linalg import spsolve

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import slab_ref  # noqa: E402
import starmesh_ref  # noqa: E402


def _coupling(system):
    """c = diag - sum of the conductances of a row (what couples the unknown to fixed potentials)"""
    rowptr, col, val, diag, rhs, mats = system
    R = diag.shape[1]
    rows = np.repeat(np.arange(R), np.diff(rowptr))
    return diag + np.stack([np.bincount(rows, weights=v, minlength=R) for v in val])


@pytest.mark.parametrize("n,seed", [(3, 0), (12, 1), (30, 2)])
def test_grid_systems(n, seed):
    system = slab_ref.random_network_system(n=n, seed=seed, nconf=3)
    rowptr, col, val, diag, rhs, mats = system
    R = diag.shape[1]
    plan = starmesh_ref.symbolic([0, R], rowptr, col)
    # every unknown is eliminated exactly once, in independent sets
    assert np.array_equal(np.sort(plan.piv), np.arange(R))
    A0 = (mats[0] != 0).astype(int)
    for r in plan.rounds[:3]:
        E = plan.piv[r["ebase"]:r["ebase"] + r["nelim"]]
        sub = A0[E][:, E]
        assert (sub - sp.diags(sub.diagonal())).nnz == 0
    x = starmesh_ref.solve(plan, val, _coupling(system), rhs)
    want = np.stack([spsolve(A.tocsc(), b) for A, b in zip(mats, rhs)])
    assert np.max(np.abs(x - want)) <= 1e-12 * 0.1


@pytest.mark.parametrize("dense_max", [4, 40, 1000])
def test_dense_tail(dense_max):
    """a network stops as soon as it has no more than dense_max unknowns left; the rest is one dense
    star-mesh elimination per network (dense_max = 1000: no sparse round at all)"""
    system = slab_ref.random_network_system(n=14, seed=6, nconf=2)
    rowptr, col, val, diag, rhs, mats = system
    R = diag.shape[1]
    # two copies of the system as a batch of two networks
    rowptr2 = np.r_[rowptr, rowptr[-1] + rowptr[1:]]
    col2 = np.r_[col, col]
    plan = starmesh_ref.symbolic([0, R, 2 * R], rowptr2, col2, dense_max=dense_max)
    assert 0 < plan.dense_nmax <= dense_max
    if dense_max >= R:
        assert len(plan.rounds) == 0
    c = _coupling(system)
    x = starmesh_ref.solve(plan, np.c_[val, val], np.c_[c, c], np.c_[rhs, rhs])
    want = np.stack([spsolve(A.tocsc(), b) for A, b in zip(mats, rhs)])
    assert np.max(np.abs(x[:, :R] - want)) <= 1e-12 * 0.1
    assert np.array_equal(x[:, :R], x[:, R:])                 # same network, same bits, wherever it sits in the batch


def test_batch_composition_does_not_change_a_networks_plan():
    """pivot order, neighbour lists and contribution order of a network are the same alone and inside a batch"""
    a = slab_ref.random_network_system(n=9, seed=3, nconf=1)
    b = slab_ref.random_network_system(n=13, seed=4, nconf=1)
    Ra, Rb = a[3].shape[1], b[3].shape[1]
    alone = starmesh_ref.symbolic([0, Rb], b[0], b[1], dense_max=20)
    both = starmesh_ref.symbolic([0, Ra, Ra + Rb], np.r_[a[0], a[0][-1] + b[0][1:]], np.r_[a[1], b[1]], dense_max=20)
    seq = lambda p, lo, hi: [(int(v) - lo, tuple(p.nbr[p.nptr[m]:p.nptr[m + 1]] - lo)) for m, v in enumerate(p.piv)
                             if lo <= v < hi]
    assert seq(alone, 0, Rb) == seq(both, Ra, Ra + Rb)
    assert np.array_equal(alone.dnodes, both.dnodes[both.dnodes >= Ra] - Ra)


def test_batch_of_networks_and_parallel_entries():
    """two disconnected networks in one pattern and duplicate entries of a pair (a virtual series
    junction next to a real one) are summed"""
    rng = np.random.default_rng(5)
    R = 0
    rows, cols, lcols, vals, roff = [], [], [], [], [0]
    for nn in (17, 9):
        # ring + chords, some pairs twice
        e = [(i, (i + 1) % nn) for i in range(nn)] + [(0, nn // 2), (1, nn // 2), (0, nn // 2)]
        for a, c in e:
            g = float(np.exp(-5 * rng.random()))
            rows += [R + a, R + c]
            cols += [R + c, R + a]
            lcols += [c, a]
            vals += [-g, -g]
        R += nn
        roff.append(R)
    order = np.lexsort((cols, rows))
    rows, cols, lcols, vals = (np.array(v)[order] for v in (rows, cols, lcols, vals))
    rowptr = np.r_[0, np.cumsum(np.bincount(rows, minlength=R))]
    c = np.zeros(R)
    b = np.zeros(R)
    c[[0, 20]] = 0.3
    b[[0, 20]] = 0.03
    c[[8, 25]] = 0.2
    A = sp.coo_matrix((vals, (rows, cols)), shape=(R, R)).tocsr()
    A = A + sp.diags(c - np.asarray(A.sum(1)).ravel())
    plan = starmesh_ref.symbolic(roff, rowptr, lcols)
    assert plan.nE0 == 17 + 2 + 9 + 2 and len(plan.e0_src) == len(vals) // 2
    x = starmesh_ref.solve(plan, vals[None], c[None], b[None])[0]
    assert np.max(np.abs(x - spsolve(A.tocsc(), b))) <= 1e-13


def test_floating_cluster_is_componentwise_accurate():
    """a strongly connected cluster hanging on the electrodes by 1e-12 contacts: the elimination has no
    cancellation, the answer is the weighted mean of the electrode potentials to machine precision
    (an LU of the assembled matrix loses ~1e-16 / 1e-12 here)"""
    n = 50
    rows, cols, vals = [], [], []
    for i in range(n - 1):
        rows += [i, i + 1]
        cols += [i + 1, i]
        vals += [-1.0, -1.0]
    rowptr = np.r_[0, np.cumsum(np.bincount(rows, minlength=n))]
    order = np.lexsort((cols, rows))
    cols, vals = np.array(cols)[order], np.array(vals)[order]
    c = np.zeros(n)
    b = np.zeros(n)
    c[0], b[0] = 1e-12, 0.1 * 1e-12          # source contact
    c[n - 1] = 3e-12                           # drain contact
    plan = starmesh_ref.symbolic([0, n], rowptr, cols)
    x = starmesh_ref.solve(plan, vals[None], c[None], b[None])[0]
    # exact: series of g_s = 1e-12, (n-1) unit conductances, g_d = 3e-12
    rs, rd = 1 / 1e-12, 1 / 3e-12
    total = rs + (n - 1) + rd
    want = 0.1 * (1 - (rs + np.arange(n)) / total)
    assert np.max(np.abs(x - want) / want) <= 1e-13


def test_complete_graphs_fill():
    """40 complete graphs: eliminating one vertex pairs up all the others -- the contribution count the
    fill budget of Engine.solve('auto') is compared with"""
    n = 12
    rows, cols = np.nonzero(~np.eye(n, dtype=bool))
    rowptr = np.r_[0, np.cumsum(np.tile(np.bincount(rows, minlength=n), 5))]
    plan = starmesh_ref.symbolic(np.arange(6) * n, rowptr, np.tile(cols, 5))
    assert plan.max_deg == n - 1
    assert len(plan.cm) == 5 * sum(d * (d - 1) // 2 for d in range(n))


def test_plan_invariants():
    """structure of the index lists every numeric kernel relies on (guards refactors of the symbolic phase)"""
    system = slab_ref.random_network_system(n=18, seed=12, nconf=1)
    rowptr, col, val, diag, rhs, mats = system
    R = diag.shape[1]
    plan = starmesh_ref.symbolic([0, R], rowptr, col, dense_max=30)
    check_plan_invariants(plan, mats[0], R, dense_max=30)


def check_plan_invariants(plan, A, R, dense_max):
    """shared with the GPU test, which runs it on the plan read back from cnt_sm_symbolic"""
    assert 0 < plan.dense_nmax <= dense_max
    A = (A != 0).astype(np.int64).tolil()
    A.setdiag(0)
    adj = {i: set(A.rows[i]) - {i} for i in range(R)}             # current graph, updated round by round
    slot_pair = {e: (int(plan.eu[e]), int(plan.ew[e])) for e in range(plan.nE0)} if hasattr(plan, "eu") else None
    seen = np.zeros(R, bool)
    dead_slots = set()
    nslots = plan.nE0
    for r in plan.rounds:
        E = plan.piv[r["ebase"]:r["ebase"] + r["nelim"]]
        assert np.all(np.diff(E) > 0) and not seen[E].any()
        seen[E] = True
        nptr = plan.nptr[r["ebase"]:r["ebase"] + r["nelim"] + 1]
        assert nptr[0] == r["lbase"] and nptr[-1] == r["lbase"] + r["nlist"]
        lo, hi = nptr[0], nptr[-1]
        nbr, eslot = plan.nbr[lo:hi], plan.eslot[lo:hi]
        for m, v in enumerate(E):
            nb = plan.nbr[nptr[m]:nptr[m + 1]]
            assert np.all(np.diff(nb) > 0) and set(nb.tolist()) == adj[int(v)]      # ascending, the true neighbourhood
            assert not (set(nb.tolist()) & set(E.tolist()))                          # independent set
        assert len(np.unique(eslot)) == len(eslot) and eslot.max(initial=-1) < nslots
        assert not (set(eslot.tolist()) & dead_slots)
        # edge targets: distinct live slots; contributions reference this round's dead edges, ordered by pivot
        gs = slice(r["gebase"], r["gebase"] + r["nge"])
        tgt = plan.tgt[gs]
        cptr = plan.cptr[r["gebase"]:r["gebase"] + r["nge"] + 1]
        assert cptr[0] == r["cbase"] and cptr[-1] == r["cbase"] + r["npairs"] and np.all(np.diff(cptr) > 0)
        assert len(np.unique(tgt)) == len(tgt) and not (set(tgt.tolist()) & (dead_slots | set(eslot.tolist())))
        fresh = tgt[tgt >= nslots]
        assert np.array_equal(fresh, nslots + np.arange(r["nnew"]))                  # fill slots numbered in group order
        inc = set(eslot.tolist())
        pos = set(range(lo, hi))                     # ca / cb / tslot: positions in this round's incident lists
        ks = slice(cptr[0], cptr[-1])
        assert set(plan.ca[ks].tolist()) <= pos and set(plan.cb[ks].tolist()) <= pos
        for k in range(r["nge"]):
            cm = plan.cm[cptr[k]:cptr[k + 1]]
            assert np.all(np.diff(cm) > 0) and cm.min() >= r["ebase"] and cm.max() < r["ebase"] + r["nelim"]
        # node targets: every neighbour once, its entries ordered by pivot
        tn = plan.tn[r["gnbase"]:r["gnbase"] + r["ngn"]]
        tptr = plan.tptr[r["gnbase"]:r["gnbase"] + r["ngn"] + 1]
        assert len(np.unique(tn)) == len(tn) and set(tn.tolist()) == set(nbr.tolist())
        assert tptr[0] == r["lbase"] and tptr[-1] == r["lbase"] + r["nlist"]
        assert sorted(plan.tslot[lo:hi].tolist()) == list(range(lo, hi))
        for k in range(r["ngn"]):
            assert np.all(np.diff(plan.tm[tptr[k]:tptr[k + 1]]) > 0)
        # update the graph: star-mesh
        for v in E:
            nb = adj.pop(int(v))
            for a in nb:
                adj[a].discard(int(v))
                adj[a] |= nb - {a}
        dead_slots |= inc
        nslots += r["nnew"]
        assert nslots - len(dead_slots) == sum(len(s) for s in adj.values()) // 2
    assert nslots == plan.nslots
    rest = plan.dnodes
    assert not seen[rest].any() and seen.sum() + len(rest) == R
    assert len(plan.dslot) == sum(len(s) for s in adj.values()) // 2 and len(np.unique(plan.depos)) == len(plan.dslot)
    assert not (set(plan.dslot.tolist()) & dead_slots)
    assert plan.nnzL == len(plan.nbr)


def test_scratch_pool_is_grow_only_and_plans_hand_their_buffer_back():
    """host logic of the direct solver's persistent work areas (starmesh.scratch): reuse while the request fits, growth
    with headroom, exclusive ownership of a taken buffer, return of a plan's buffer when the plan dies"""
    import torch
    from networksim_cntfet_b200 import starmesh
    pool = {}
    a = starmesh.scratch(pool, "ws", 1000, "cpu")
    assert a.numel() >= 1000 and starmesh.scratch(pool, "ws", 900, "cpu") is a          # fits: same buffer
    b = starmesh.scratch(pool, "ws", a.numel() + 1, "cpu")
    assert b is not a and b.numel() >= (a.numel() + 1) * 8 // 7 and pool["ws"] is b      # grew with headroom
    assert starmesh.scratch(None, "ws", 10, "cpu").numel() == 10                         # no pool: exact size
    p = starmesh.Plan()
    p.buf = starmesh.scratch(pool, "sm_plan", 5000, "cpu", take=True)
    p._pool = pool
    assert pool["sm_plan"] is None                                                        # the plan owns it
    q = starmesh.scratch(pool, "sm_plan", 5000, "cpu", take=True)                         # a second live plan gets its own
    assert q is not p.buf
    buf = p.buf
    del p
    assert pool["sm_plan"] is buf                                                         # handed back
    assert starmesh.scratch(pool, "sm_plan", 4000, "cpu", take=True) is buf and pool["sm_plan"] is None
