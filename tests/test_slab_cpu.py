"""Sharded solve of one network (networksim_cntfet_b200/slab.py): partition plan + exchange schedule
on CPU with the torch restatement of the phases (tests/slab_ref.py) -- in-process ranks and a
world-size-2 gloo job -- against scipy's direct solve."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import slab_ref  # noqa: E402
from networksim_cntfet_b200 import slab  # noqa: E402


def _solve_direct(mats, rhs):
    from scipy.sparse.linalg import spsolve
    return np.stack([spsolve(A.tocsc(), b) for A, b in zip(mats, rhs)])


@pytest.mark.parametrize("world", [1, 2, 3, 5])
@pytest.mark.parametrize("precond", [True, False])
def test_inprocess_ranks_match_direct_solve(world, precond):
    system = slab_ref.random_network_system(n=24, seed=world, nconf=2)
    comm = slab.LocalComm(world)
    ph = slab_ref.TorchPhases()
    states = slab_ref.to_states(system, comm, ph, precond=precond)
    if world > 1:
        assert all(st.n_halo > 0 and st.n_exp > 0 for st in states)
        if precond:
            assert int(states[0].nsh.max()) > 0           # some aggregates span ranks
    issued = slab.run_pcg(states, comm, ph, check_every=20)
    assert all(bool(st.done.all()) for st in states) and issued < 20000
    assert all(float(st.scal[:, slab_ref.STATUS].abs().max()) == 0.0 for st in states)
    R = system[3].shape[1]
    x = slab.gather_solution(states, comm, R).numpy()
    want = _solve_direct(system[5], system[4])
    assert np.max(np.abs(x - want)) <= 1e-10 * 0.1
    # every rank saw the same scalars
    for st in states[1:]:
        assert torch.equal(st.scal, states[0].scal)


def test_more_ranks_than_rows():
    """ranks that own no row at all take part in the exchanges with empty slices"""
    system = slab_ref.random_network_system(n=3, seed=1, nconf=2)          # 9 unknowns
    comm = slab.LocalComm(12)
    ph = slab_ref.TorchPhases()
    states = slab_ref.to_states(system, comm, ph)
    assert sum(st.n_own == 0 for st in states) >= 3
    slab.run_pcg(states, comm, ph, check_every=5)
    x = slab.gather_solution(states, comm, 9).numpy()
    assert np.max(np.abs(x - _solve_direct(system[5], system[4]))) <= 1e-10 * 0.1


def test_rank_count_does_not_change_iteration_count_much():
    """the preconditioner is the same operator for any number of ranks (shared aggregates are
    reduced exactly), so the iteration counts agree up to rounding effects"""
    system = slab_ref.random_network_system(n=24, seed=7, nconf=2)
    its = []
    for world in (1, 4):
        comm = slab.LocalComm(world)
        ph = slab_ref.TorchPhases()
        states = slab_ref.to_states(system, comm, ph)
        slab.run_pcg(states, comm, ph, check_every=10)
        its.append(states[0].scal[:, slab_ref.ITERS].clone())
    assert torch.all((its[0] - its[1]).abs() <= 2), its


def _gloo_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        system = slab_ref.random_network_system(n=20, seed=3, nconf=2)
        comm = slab.DistComm()
        ph = slab_ref.TorchPhases()
        states = slab_ref.to_states(system, comm, ph)
        slab.run_pcg(states, comm, ph, check_every=25)
        x = slab.gather_solution(states, comm, system[3].shape[1]).numpy()
        want = _solve_direct(system[5], system[4])
        ok = bool(np.max(np.abs(x - want)) <= 1e-10 * 0.1) and bool(states[0].done.all())
        open(os.path.join(out, "rank%d" % rank), "w").write("ok" if ok else "bad %g" % np.max(np.abs(x - want)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert [open(os.path.join(str(tmp_path), "rank%d" % r)).read() for r in range(2)] == ["ok", "ok"]
