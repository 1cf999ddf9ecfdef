"""GPU: ONE network with its unknown rows sharded over several ranks (Engine.solve_sharded,
networksim_cntfet_b200/slab.py, csrc/pcg_slab.cu).  slab.LocalComm runs all W ranks of a plan on
this GPU -- same kernels, plans and exchange schedule as the NCCL job, the collectives being
copies -- so the multi-rank logic is covered on a single device; tools/big_network_dist.py
(BIG_MODE=rows) is the same thing over torch.distributed / NCCL."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import slab_ref  # noqa: E402
from networksim_cntfet_b200 import slab  # noqa: E402
from oracle import cntnet_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu
VDS = 0.1


def _direct(system):
    from scipy.sparse.linalg import spsolve
    return np.stack([spsolve(A.tocsc(), b) for A, b in zip(system[5], system[4])])


@pytest.mark.parametrize("world", [1, 3, 8])
@pytest.mark.parametrize("precond", [True, False])
def test_cuda_phases_match_direct_solve(engine, world, precond):
    system = slab_ref.random_network_system(n=40, seed=world, nconf=3)
    comm = slab.LocalComm(world)
    with engine._work():
        ph = slab.CudaPhases(engine.stream_ptr)
        states = slab_ref.to_states(system, comm, ph, device=engine.device, precond=precond)
        issued = slab.run_pcg(states, comm, ph, check_every=20)
        x = slab.gather_solution(states, comm, system[3].shape[1])
    engine.sync()
    assert all(bool(st.done.all()) for st in states) and issued < 20000
    assert float(states[0].scal[:, slab_ref.STATUS].abs().max()) == 0.0
    assert np.max(np.abs(x.cpu().numpy() - _direct(system))) <= 1e-10 * VDS
    for st in states[1:]:
        assert torch.equal(st.scal, states[0].scal)


def test_more_ranks_than_rows(engine):
    """ranks that own no row at all take part in the exchanges with empty slices"""
    system = slab_ref.random_network_system(n=3, seed=1, nconf=2)          # 9 unknowns
    comm = slab.LocalComm(12)
    with engine._work():
        ph = slab.CudaPhases(engine.stream_ptr)
        states = slab_ref.to_states(system, comm, ph, device=engine.device)
        slab.run_pcg(states, comm, ph, check_every=5)
        x = slab.gather_solution(states, comm, 9)
    engine.sync()
    assert sum(st.n_own == 0 for st in states) >= 3
    assert np.max(np.abs(x.cpu().numpy() - _direct(system))) <= 1e-10 * VDS


def test_cuda_phases_step_by_step_against_torch(engine):
    """every phase leaves the same numbers in the shared state as the torch restatement"""
    system = slab_ref.random_network_system(n=32, seed=11, nconf=2)
    W = 4
    cg, cc = slab.LocalComm(W), slab.LocalComm(W)
    tp = slab_ref.TorchPhases()
    with engine._work():
        gp = slab.CudaPhases(engine.stream_ptr)
        sg = slab_ref.to_states(system, cg, gp, device=engine.device)
    sc = slab_ref.to_states(system, cc, tp, device="cpu")

    def same(tag):
        engine.sync()
        for a, b in zip(sg, sc):
            for name in ("x", "r", "p", "S", "red2", "red_pq", "scal", "sendbuf"):
                u, v = getattr(a, name).cpu(), getattr(b, name)
                scale = float(v.abs().max()) or 1.0
                assert float((u - v).abs().max()) <= 1e-9 * scale, (tag, a.rank, name)    # FMA contraction differs
            assert torch.equal(a.done.cpu(), b.done)

    def both(fn):
        with engine._work():
            fn(sg, cg, gp)
        fn(sc, cc, tp)

    both(lambda S, c, p: [p.init(st) for st in S])
    same("init")
    both(lambda S, c, p: c.all_reduce([st.red2 for st in S]))
    both(lambda S, c, p: [p.direction(st, 1) for st in S])
    same("direction0")
    both(lambda S, c, p: c.all_gather([st.recvbuf for st in S], [st.sendbuf for st in S]))
    for it in range(5):
        both(lambda S, c, p: [p.spmv(st) for st in S])
        same("spmv%d" % it)
        both(lambda S, c, p: c.all_reduce([st.red_pq for st in S]))
        both(lambda S, c, p: [p.update(st) for st in S])
        same("update%d" % it)
        both(lambda S, c, p: c.all_reduce([st.red2 for st in S]))
        both(lambda S, c, p: [p.direction(st, 0) for st in S])
        same("direction%d" % it)
        both(lambda S, c, p: c.all_gather([st.recvbuf for st in S], [st.sendbuf for st in S]))


@pytest.mark.parametrize("world", [2, 5])
def test_sharded_fullnet_matches_oracle_and_unsharded(engine, world):
    """20x20 um network: measure_fullnet with the rows sharded over `world` ranks gives the oracle's
    currents (1e-9) and the unsharded engine's voltages"""
    b = engine.generate([6400], 20.0, seed=5)
    ref = engine.measure_fullnet(b, onoffmap=1, want_fields=True)
    b2 = engine.generate([6400], 20.0, seed=5)
    res = engine.measure_fullnet(b2, onoffmap=1, want_fields=True, row_shard=slab.LocalComm(world))
    assert engine.last_solver == "slab x%d" % world
    assert res["percolating"][0] and np.all(res["status"] == 0)
    m = O.measure_fullnet(engine.sticks_to_host(b2)[0], onoffmap=1, refine=3)
    for key in ("ion", "ioff_back", "ioff_total", "ioff_partial"):
        assert abs(res[key][0] - m[key]) <= 1e-9 * m[key], (key, res[key][0], m[key])
        assert abs(res[key][0] - ref[key][0]) <= 1e-9 * ref[key][0]
    V, Vr = res["fields"]["V"].cpu().numpy(), ref["fields"]["V"].cpu().numpy()
    ok = ~np.isnan(Vr)
    assert np.array_equal(np.isnan(V), ~ok)
    assert np.max(np.abs(V[ok] - Vr[ok])) <= 1e-9 * VDS
    # the preconditioner does not depend on the number of ranks: iteration counts as on one rank
    one = engine.measure_fullnet(engine.generate([6400], 20.0, seed=5), onoffmap=1, row_shard=slab.LocalComm(1))
    assert np.all(np.abs(one["iters"].astype(int) - res["iters"].astype(int)) <= 0.05 * one["iters"] + 5)


def test_sharded_full_size_100um(engine):
    """100x100 um (46k unknowns after the exact reductions) on 8 ranks vs the cluster solver"""
    engine.direct_auto = False          # reference: the iterative cluster solver
    try:
        ref = engine.measure_fullnet(engine.generate([100000], 100.0, seed=9), onoffmap=1)
        assert engine.last_solver.startswith("cluster")
    finally:
        engine.direct_auto = True
    res = engine.measure_fullnet(engine.generate([100000], 100.0, seed=9), onoffmap=1, row_shard=slab.LocalComm(8))
    assert np.all(res["status"] == 0)
    states = engine.last_slab_states
    assert len(states) == 8 and all(st.n_halo > 0 for st in states) and int(states[0].nsh.max()) > 0
    for key in ("ion", "ioff_back", "ioff_total", "ioff_partial"):
        assert abs(res[key][0] - ref[key][0]) <= 1e-9 * ref[key][0], (key, res[key][0], ref[key][0])


def test_large_single_network_takes_slab_solver(engine):
    """a single network beyond the cluster kernel's capacity (160x160 um, ~118k unknowns) is solved by
    the row-slab solver with one rank; currents against the refined oracle"""
    b = engine.generate([256000], 160.0, seed=3)
    engine.direct_auto = False          # the iterative route (the direct solver would take it otherwise)
    try:
        res = engine.measure_fullnet(b, onoffmap=1)
    finally:
        engine.direct_auto = True
    from networksim_cntfet_b200 import _lib
    assert b.R > _lib.load().cnt_pcg_cluster_max_rows() and engine.last_solver == "slab x1"
    assert np.all(res["status"] == 0)
    m = O.measure_fullnet(engine.sticks_to_host(b)[0], onoffmap=1, refine=3)
    for key in ("ion", "ioff_back", "ioff_total", "ioff_partial"):
        assert abs(res[key][0] - m[key]) <= 1e-9 * m[key], (key, res[key][0], m[key])


def test_graph_replay_equals_eager(engine):
    """blocks of iterations replayed as a CUDA graph (kernels + the in-process collectives) give the
    bits of the eager schedule"""
    out = []
    for graph in (False, True):
        engine.slab_graph = graph
        try:
            res = engine.measure_fullnet(engine.generate([6400], 20.0, seed=5), onoffmap=1, want_fields=True,
                                         row_shard=slab.LocalComm(3))
        finally:
            engine.slab_graph = True
        assert engine.last_slab_info["graph"] == graph
        out.append(res)
    assert np.array_equal(out[0]["iters"], out[1]["iters"])
    assert torch.equal(out[0]["fields"]["x"], out[1]["fields"]["x"])


def test_non_percolating_network_sharded(engine):
    b = engine.generate([120], 5.0, seed=5)
    res = engine.measure_fullnet(b, onoffmap=1, row_shard=slab.LocalComm(3))
    if not res["percolating"][0]:
        assert res["ion"][0] == 0.0 and np.all(res["iters"] == 0)
