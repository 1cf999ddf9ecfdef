"""CPU: symbolic phase of the star-mesh direct solver (networksim_cntfet_b200/starmesh.py) with the
torch restatement of its numeric phase (tests/starmesh_ref.py) against scipy's direct solve."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp
import torch
from scipy.sparse.linalg import spsolve

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import slab_ref  # noqa: E402
import starmesh_ref  # noqa: E402
from networksim_cntfet_b200 import starmesh  # noqa: E402


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
    t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt)
    plan = starmesh.symbolic(t(rowptr, torch.int64), t(col, torch.int64), R)
    # every unknown is eliminated exactly once, in independent sets
    E = torch.cat([r.E for r in plan.rounds])
    assert np.array_equal(np.sort(E.numpy()), np.arange(R))
    A0 = (mats[0] != 0).astype(int)
    for r in plan.rounds[:3]:
        sub = A0[r.E.numpy()][:, r.E.numpy()]
        assert (sub - sp.diags(sub.diagonal())).nnz == 0
    x = starmesh_ref.solve(plan, t(val, torch.float64), t(_coupling(system), torch.float64), t(rhs, torch.float64)).numpy()
    want = np.stack([spsolve(A.tocsc(), b) for A, b in zip(mats, rhs)])
    assert np.max(np.abs(x - want)) <= 1e-12 * 0.1


@pytest.mark.parametrize("dense_max", [4, 40, 1000])
def test_dense_tail(dense_max):
    """the rounds stop when no network has more than dense_max unknowns left; the rest is one dense
    star-mesh elimination per network (dense_max = 1000: no sparse round at all)"""
    system = slab_ref.random_network_system(n=14, seed=6, nconf=2)
    rowptr, col, val, diag, rhs, mats = system
    R = diag.shape[1]
    # two copies of the system as a batch of two networks
    rowptr2 = np.r_[rowptr, rowptr[-1] + rowptr[1:]]
    col2 = np.r_[col, col + R]
    t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt)
    net_ids = t(np.repeat([0, 1], R), torch.int64)
    local = t(np.r_[np.arange(R), np.arange(R)], torch.int64)
    plan = starmesh.symbolic(t(rowptr2, torch.int64), t(col2, torch.int64), 2 * R, local_ids=local, net_ids=net_ids, nnet=2,
                             dense_max=dense_max)
    assert plan.dense is not None and plan.dense["nmax"] <= dense_max
    if dense_max >= R:
        assert len(plan.rounds) == 0
    c = _coupling(system)
    x = starmesh_ref.solve(plan, t(np.c_[val, val], torch.float64), t(np.c_[c, c], torch.float64),
                           t(np.c_[rhs, rhs], torch.float64)).numpy()
    want = np.stack([spsolve(A.tocsc(), b) for A, b in zip(mats, rhs)])
    assert np.max(np.abs(x[:, :R] - want)) <= 1e-12 * 0.1
    assert np.array_equal(x[:, :R], x[:, R:])                 # same network, same bits, wherever it sits in the batch
    plan = starmesh.finalize(plan, keep64=True)
    assert plan.dense["dptr32"].dtype == torch.int32


def test_batch_of_networks_and_parallel_entries():
    """two disconnected networks in one pattern (global indices) and duplicate entries of a pair
    (a virtual series junction next to a real one) are summed"""
    rng = np.random.default_rng(5)
    blocks, R = [], 0
    rows, cols, vals = [], [], []
    for nn in (17, 9):
        # ring + chords, some pairs twice
        e = [(i, (i + 1) % nn) for i in range(nn)] + [(0, nn // 2), (1, nn // 2), (0, nn // 2)]
        for a, c in e:
            g = float(np.exp(-5 * rng.random()))
            rows += [R + a, R + c]
            cols += [R + c, R + a]
            vals += [-g, -g]
        R += nn
    order = np.lexsort((cols, rows))
    rows, cols, vals = np.array(rows)[order], np.array(cols)[order], np.array(vals)[order]
    rowptr = np.r_[0, np.cumsum(np.bincount(rows, minlength=R))]
    c = np.zeros(R)
    b = np.zeros(R)
    c[[0, 20]] = 0.3
    b[[0, 20]] = 0.03
    c[[8, 25]] = 0.2
    A = sp.coo_matrix((vals, (rows, cols)), shape=(R, R)).tocsr()
    A = A + sp.diags(c - np.asarray(A.sum(1)).ravel())
    t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt)
    plan = starmesh.symbolic(t(rowptr, torch.int64), t(cols, torch.int64), R)
    x = starmesh_ref.solve(plan, t(vals[None], torch.float64), t(c[None], torch.float64), t(b[None], torch.float64)).numpy()[0]
    assert np.max(np.abs(x - spsolve(A.tocsc(), b))) <= 1e-13
    plan = starmesh.finalize(plan, keep64=True)
    assert plan.pool.dtype == torch.int32 and plan.nrounds == len(plan.rounds) and plan.nE_max >= plan.nE0


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
    t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt)
    plan = starmesh.symbolic(t(rowptr, torch.int64), t(cols, torch.int64), n)
    x = starmesh_ref.solve(plan, t(vals[None], torch.float64), t(c[None], torch.float64), t(b[None], torch.float64)).numpy()[0]
    # exact: series of g_s = 1e-12, (n-1) unit conductances, g_d = 3e-12
    rs, rd = 1 / 1e-12, 1 / 3e-12
    total = rs + (n - 1) + rd
    want = 0.1 * (1 - (rs + np.arange(n)) / total)
    assert np.max(np.abs(x - want) / want) <= 1e-13


def test_fill_budget_rejects_dense_graphs():
    """a dense, mesh-like pattern exceeds the contribution budget: symbolic() returns None and the
    engine falls back to the iterative solvers (Engine.solve, 'auto')"""
    n = 60
    rows, cols = np.nonzero(~np.eye(n, dtype=bool))
    rowptr = np.r_[0, np.cumsum(np.bincount(rows, minlength=n))]
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.int64)
    # complete graph: eliminating one vertex pairs up all the others; far more than 100000 + 0.5 per edge
    # only for big graphs, so shrink the constant part by scaling the problem: 40 disjoint copies
    R = 40 * n
    rowptr_b = np.r_[0, np.cumsum(np.tile(np.diff(rowptr), 40))]
    cols_b = np.concatenate([cols + k * n for k in range(40)])
    assert starmesh.symbolic(t(rowptr_b), t(cols_b), R, max_fill=0.5) is None
    plan = starmesh.symbolic(t(rowptr_b), t(cols_b), R, max_fill=None)
    assert plan is not None and plan.max_deg == n - 1


def test_plan_invariants():
    """structure of the index lists every numeric kernel relies on (guards refactors of the symbolic phase)"""
    system = slab_ref.random_network_system(n=18, seed=12, nconf=1)
    rowptr, col, val, diag, rhs, mats = system
    R = diag.shape[1]
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.int64)
    plan = starmesh.symbolic(t(rowptr), t(col), R, net_ids=torch.zeros(R, dtype=torch.int64), nnet=1, dense_max=30)
    assert plan.dense is not None and 0 < plan.dense["nmax"] <= 30
    A = (mats[0] != 0).astype(np.int64).tolil()
    A.setdiag(0)
    adj = {i: set(A.rows[i]) - {i} for i in range(R)}             # current graph, updated round by round
    seen = np.zeros(R, bool)
    gbase = 0
    for r in plan.rounds:
        E = r.E.numpy()
        assert np.all(np.diff(E) > 0) and not seen[E].any()
        seen[E] = True
        nptr, nbr, eslot = r.nptr.numpy(), r.nbr.numpy(), r.eslot.numpy()
        assert nptr[0] == 0 and nptr[-1] == len(nbr) and r.gbase == gbase
        gbase += len(nbr)
        for m, v in enumerate(E):
            nb = nbr[nptr[m]:nptr[m + 1]]
            assert np.all(np.diff(nb) > 0) and set(nb.tolist()) == adj[int(v)]      # ascending, the true neighbourhood
            assert not (set(nb.tolist()) & set(E.tolist()))                          # independent set
        assert len(np.unique(eslot)) == len(eslot) and eslot.max(initial=-1) < r.nE_old
        # new edge list: every surviving edge copied exactly once, every pair contribution lands on an edge
        cs = r.copy_src.numpy()
        kept = cs[cs >= 0]
        assert len(np.unique(kept)) == len(kept) and not (set(kept.tolist()) & set(eslot.tolist()))
        assert len(kept) + len(eslot) == r.nE_old
        cptr = r.cptr.numpy()
        assert cptr[0] == 0 and cptr[-1] == r.cm.numel() and np.all(np.diff(cptr) >= 0) and len(cptr) == r.nE_new + 1
        assert np.all((cs >= 0) | (np.diff(cptr) > 0))                               # a new edge has at least one contribution
        inc = set(eslot.tolist())
        assert set(r.ca.numpy().tolist()) <= inc and set(r.cb.numpy().tolist()) <= inc
        tn = r.tn.numpy()
        assert np.all(np.diff(tn) > 0) and set(tn.tolist()) == set(nbr.tolist())
        assert r.tptr.numpy()[-1] == len(nbr) and sorted(r.tslot.numpy().tolist()) == sorted(eslot.tolist())
        # update the graph: star-mesh
        for v in E:
            nb = adj.pop(int(v))
            for a in nb:
                adj[a].discard(int(v))
                adj[a] |= nb - {a}
        assert r.nE_new == sum(len(s) for s in adj.values()) // 2
    d = plan.dense
    rest = d["dnodes"].numpy()
    assert not seen[rest].any() and seen.sum() + len(rest) == R
    assert d["nE"] == sum(len(s) for s in adj.values()) // 2 and len(np.unique(d["epos"].numpy())) == d["nE"]
    assert plan.nnzL == gbase
