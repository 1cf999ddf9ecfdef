"""CPU: host logic of the slab-sharded junction search (Engine._find_junctions_slab): the slab partition of the
tile columns and the exchange of per-rank tables with disjoint supports -- in-process ranks and a world-size-2
gloo job.  The CUDA side is covered by tests/test_gpu_slab_junctions.py."""
import os
import socket
import sys

import numpy as np
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from networksim_cntfet_b200 import slab  # noqa: E402


def test_slab_partition_is_contiguous_balanced_and_complete():
    for TG in (1, 2, 7, 28, 280):
        for world in (1, 2, 3, 8, 16):
            own = [slab.junction_slab_owner(c, TG, world) for c in range(TG)]
            assert own == sorted(own) and 0 <= min(own) and max(own) < world       # contiguous slabs, left to right
            counts = np.bincount(own, minlength=world)
            assert counts.sum() == TG and counts.max() - counts.min() <= 1          # every column once, balanced


def _tables(world, n=1000, seed=3):
    """a full junction table and its split into per-rank tables with disjoint supports (zeros elsewhere)"""
    rng = np.random.default_rng(seed)
    full = {"stick1": rng.integers(0, 1 << 20, n).astype(np.int32), "stick2": rng.integers(0, 1 << 20, n).astype(np.int32),
            "jx": rng.random(n), "jy": rng.random(n), "kind2": rng.integers(0, 9, n).astype(np.uint8)}
    full["jx"][:4] = [-0.0, np.inf, 5e-324, 1.0]                    # bit patterns that a floating-point sum would not keep
    full["jy"][:2] = [np.float64(np.nan), -0.0]
    owner = rng.integers(0, world, n)
    parts = [{k: np.where(owner == r, v, np.zeros(1, dtype=v.dtype)) for k, v in full.items()} for r in range(world)]
    return full, parts


def _same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))


def test_sum_disjoint_inprocess_ranks():
    for world in (2, 3, 8):
        full, parts = _tables(world)
        comm = slab.LocalComm(world)
        for key in full:
            ts = [torch.from_numpy(p[key].copy()) for p in parts]
            slab.sum_disjoint(comm, ts)
            for t in ts:
                assert _same_bits(t.numpy(), full[key]), (world, key)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full, parts = _tables(world)
        comm = slab.DistComm()
        ok = True
        for key in full:
            t = torch.from_numpy(parts[rank][key].copy())
            slab.sum_disjoint(comm, [t])
            ok = ok and _same_bits(t.numpy(), full[key])
        q.put((rank, ok, comm.world))
    finally:
        dist.destroy_process_group()


def test_sum_disjoint_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == [(0, True, 2), (1, True, 2)]
