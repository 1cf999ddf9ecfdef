"""CPU, build container only: the numpy restatement (oracle/cntnet_oracle.py) against the UNMODIFIED reference
imported through oracle/refshim.py, on fresh seeds (the committed goldens pin the same thing from files; this
test re-derives it live and is skipped wherever /root/reference does not exist, e.g. on the GPU box)."""
import contextlib
import io

import numpy as np
import pytest

from oracle import cntnet_oracle as O
from oracle import refshim

pytestmark = pytest.mark.skipif(not refshim.available(), reason="reference not present (build container only)")

KCODE = {"v": 0, "m": 1, "s": 2}


@pytest.mark.parametrize("n,scaling,seed,onoffmap", [(250, 5, 11, 0), (300, 5, 12, 1), (280, 5, 13, 2), (150, 5, 14, 0)])
def test_oracle_equals_reference(n, scaling, seed, onoffmap):
    netsim, cnet = refshim.load(hoist=True)
    with contextlib.redirect_stdout(io.StringIO()):
        net = netsim.RandomCNTNetwork(n=n, scaling=scaling, seed=seed, onoffmap=onoffmap)
    st = net.sticks
    ends = np.stack(st.endarray.values)
    s = dict(xc=st.xc.values.astype(float), yc=st.yc.values.astype(float), length=st.length.values.astype(float),
             xa=ends[:, 0, 0].copy(), ya=ends[:, 0, 1].copy(), xb=ends[:, 1, 0].copy(), yb=ends[:, 1, 1].copy(),
             kind=np.array([KCODE[k] for k in st.kind.values], dtype=np.uint8), orig=st.cluster.values.astype(np.int64))
    it = net.intersects
    order = np.lexsort((it.stick2.values, it.stick1.values))
    j = O.find_junctions(s)
    assert np.array_equal(j["stick1"], it.stick1.values[order]) and np.array_equal(j["stick2"], it.stick2.values[order])
    assert np.array_equal(j["x"], it.x.values[order]) and np.array_equal(j["y"], it.y.values[order])      # bit-exact
    c = O.label_clusters(len(st), j["stick1"], j["stick2"], s["orig"])
    assert c["percolating"] == bool(net.percolating)
    if not net.percolating:
        return
    m = O.measure_fullnet(s, onoffmap=onoffmap, junctions=j)
    ion = float(sum(net.cnet.source_currents))
    assert abs(m["ion"] - ion) <= 1e-9 * ion
    net.cnet.set_global_gate(10)
    net.cnet.update()
    ioff = float(sum(net.cnet.source_currents))
    assert abs(m["ioff_back"] - ioff) <= 5e-9 * ioff        # SuperLU at Vg = +10: ~2e-9
