"""GPU: the junction search sharded by spatial slab (cnt_junction_count_slab / _fill_slab, Engine._find_junctions_slab)
with W in-process ranks on one device (slab.LocalComm: the collectives are copies, the kernels and the exchange
schedule are those of the multi-GPU run): tables, counts and everything downstream are bit-identical to the unsharded
search."""
import numpy as np
import pytest
import torch

from conftest import golden_names, golden_sticks, load_golden
from networksim_cntfet_b200 import slab

pytestmark = pytest.mark.gpu
KEYS = ("stick1", "stick2", "jx", "jy", "kind2", "joff", "ntests", "jstart")


def _tables(b):
    return {k: getattr(b, k).clone() for k in KEYS}


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_search_equals_unsharded_on_a_ragged_batch(engine, world):
    gs = [load_golden(n) for n in golden_names()]
    mk = lambda: engine.batch_from_host([golden_sticks(g) for g in gs])
    ref = engine.find_junctions(mk())
    got = engine.find_junctions(mk(), comm=slab.LocalComm(world))
    assert got.J == ref.J
    for k in KEYS:
        a, c = getattr(ref, k), getattr(got, k)
        n = ref.J if k in ("stick1", "stick2", "jx", "jy", "kind2") else a.numel()
        assert torch.equal(a[:n], c[:n]), (world, k)


@pytest.mark.parametrize("world,n,side", [(8, 100000, 100.0), (5, 6400, 20.0), (16, 900000, 300.0)])
def test_sharded_search_full_size_and_downstream(engine, world, n, side):
    """one network: every rank's slab gets work (tile columns >> ranks), the electrodes are rank 0's; labels,
    currents of the four measure_fullnet gates identical to the unsharded run"""
    ref = engine.generate([n], side, seed=11)
    got = engine.generate([n], side, seed=11)
    r0 = engine.measure_fullnet(engine.prepare(ref), onoffmap=1)
    r1 = engine.measure_fullnet(engine.prepare(got, slab.LocalComm(world)), onoffmap=1)
    assert got.J == ref.J and ref.J > n
    for k in ("stick1", "stick2", "jx", "jy", "kind2"):
        assert torch.equal(getattr(ref, k)[:ref.J], getattr(got, k)[:ref.J]), k
    assert torch.equal(ref.ntests, got.ntests) and torch.equal(ref.label, got.label)
    for k in ("ion", "ioff_back", "ioff_total", "ioff_partial", "nclust", "maxclust"):
        assert np.array_equal(r0[k], r1[k]), k
