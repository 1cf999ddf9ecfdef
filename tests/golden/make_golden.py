#!/usr/bin/env python3
"""Generate the committed golden fixtures by running the UNMODIFIED reference
(`/root/reference`, through `oracle/refshim.py`) in the build container.

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

Each fixture holds, for one network: the reference's POST-sort stick table
(`netsim.py:131-134`), its junction list sorted by (stick1, stick2), min-index cluster labels
derived from `nx.connected_components`, the reference's own nclust / maxclust
(`measure_perc.py:76-80`), and -- when percolating -- node voltages, |edge currents| and
source currents for the Vg=0 solve done in the constructor, for the `measure_fullnet` gate
sequence (`measure_perc.py:166-186`) and for a `RandomCNTNetwork.gate` sweep
(`measure_perc.py:39-53`).  The GPU box never runs this script (no reference there).
"""
import io
import os
import sys
import contextlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refshim  # noqa: E402

KCODE = {"v": 0, "m": 1, "s": 2}


def snapshot_solution(net):
    """V per stick (nan outside the component), |I| per junction row of the sorted table."""
    g = net.cnet.graph
    n = len(net.sticks)
    V = np.full(n, np.nan)
    for node in g.nodes():
        V[node] = g.nodes[node]["voltage"]
    return V, float(sum(net.cnet.source_currents))


def edge_currents(net, st1, st2):
    g = net.cnet.graph
    I = np.full(len(st1), np.nan)
    G = np.full(len(st1), np.nan)
    for k, (a, b) in enumerate(zip(st1, st2)):
        if g.has_edge(a, b):
            I[k] = g.edges[a, b]["current"]
            G[k] = g.edges[a, b]["conductance"]
    return I, G


def build(name, cls="RandomCNTNetwork", trivial=False, element="LinExpTransistor", sweep=True, **kw):
    netsim, cnet = refshim.load(hoist=True)
    import networkx as nx
    elem = getattr(cnet, element)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):       # non-percolating nets print a traceback (netsim.py:215-218)
        if trivial:
            net = netsim.RandomCNTNetwork(n=2, scaling=1, seed=1, onoffmap=kw.get("onoffmap", 0), element=elem)
            net.percolating = False
            net.make_trivial_sticks()
        else:
            net = getattr(netsim, cls)(element=elem, **kw)
    st = net.sticks
    n = len(st)
    ends = np.stack(st.endarray.values)            # [n,2,2]
    out = dict(
        xc=st.xc.values.astype(np.float64), yc=st.yc.values.astype(np.float64),
        angle=st.angle.values.astype(np.float64), length=st.length.values.astype(np.float64),
        kind=np.array([KCODE[k] for k in st.kind.values], dtype=np.uint8),
        xa=ends[:, 0, 0].copy(), ya=ends[:, 0, 1].copy(), xb=ends[:, 1, 0].copy(), yb=ends[:, 1, 1].copy(),
        orig=st.cluster.values.astype(np.int64),    # pre-sort index (netsim.py:132), before label_clusters
        scaling=np.float64(net.scaling), seed=np.int64(net.seed), onoffmap=np.int64(net.onoffmap),
        element=np.array(element), percolating=np.bool_(net.percolating),
    )
    it = net.intersects
    order = np.lexsort((it.stick2.values, it.stick1.values))
    st1 = it.stick1.values[order].astype(np.int32)
    st2 = it.stick2.values[order].astype(np.int32)
    out.update(stick1=st1, stick2=st2, jx=it.x.values[order].astype(np.float64),
               jy=it.y.values[order].astype(np.float64),
               kind2=np.array([3 * KCODE[k[0]] + KCODE[k[1]] for k in it.kind.values[order]], dtype=np.uint8))
    # min-index labels from the reference's own component finder
    label = np.arange(n, dtype=np.int32)
    full_graph = nx.from_pandas_edgelist(net.intersects, source="stick1", target="stick2")
    for c in nx.connected_components(full_graph):
        m = min(c)
        for v in c:
            label[v] = m
    out["label"] = label
    # reference cluster statistics exactly as measure_perc.py:75-80 computes them
    ref_graph = net.graph if not isinstance(net.graph, tuple) else None
    with contextlib.redirect_stdout(buf):
        net.label_clusters()
    out["nclust_ref"] = np.int64(len(net.sticks.cluster.drop_duplicates()))
    try:
        out["maxclust_ref"] = np.int64(len(max(nx.connected_components(net.graph))))
    except Exception:
        out["maxclust_ref"] = np.int64(0)
    out["clustersizes"] = np.asarray(net.clustersizes, dtype=np.int64)

    if net.percolating:
        V, isrc = snapshot_solution(net)
        I, G = edge_currents(net, st1, st2)
        out.update(V0=V, I0=I, G0=G, isrc0=np.float64(isrc))
        # measure_fullnet gate sequence (measure_perc.py:166-186)
        c = net.cnet
        seq = {}
        c.set_global_gate(10); c.update(); seq["back"] = snapshot_solution(net)
        Ib, Gb = edge_currents(net, st1, st2)
        c.set_global_gate(0); c.set_local_gate([0.217, 0.5, 0.167, 1.2], 10); c.update()
        seq["total"] = snapshot_solution(net)
        It, Gt = edge_currents(net, st1, st2)
        c.set_global_gate(0); c.set_local_gate([0.5, 0, 0.16, 0.667], 10); c.update()
        seq["partial"] = snapshot_solution(net)
        Ip, Gp = edge_currents(net, st1, st2)
        out.update(V_back=seq["back"][0], isrc_back=np.float64(seq["back"][1]), I_back=Ib, G_back=Gb,
                   V_total=seq["total"][0], isrc_total=np.float64(seq["total"][1]), I_total=It, G_total=Gt,
                   V_partial=seq["partial"][0], isrc_partial=np.float64(seq["partial"][1]), I_partial=Ip, G_partial=Gp)
        if sweep and hasattr(net, "gate"):
            vgs = np.linspace(-10, 10, 3)
            cur = []
            for gname in ("back", "partial", "total"):       # add_voltagemeas order, measure_perc.py:44
                for vg in vgs:
                    cur.append(net.gate(vg, gname))
            out.update(sweep_vg=vgs, sweep_current=np.array(cur, dtype=np.float64))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("%-28s n=%6d J=%6d perc=%s isrc0=%s nclust_ref=%d maxclust_ref=%d" % (
        name, n, len(st1), bool(net.percolating), out.get("isrc0"), out["nclust_ref"], out["maxclust_ref"]))


def dump_reference_system(name, **kw):
    """`<name>_sticks.csv` / `<name>_intersects.csv` written by the REFERENCE's own save_system (netsim.py:228-236),
    plus the reference's currents for that network -- the N2 parity input: the GPU path loads the reference's dump
    through load_system (tests/test_gpu_api.py::test_load_reference_dump)"""
    netsim, cnet = refshim.load(hoist=True)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        net = netsim.RandomCNTNetwork(**kw)
    net.save_system(os.path.join(HERE, name))
    V, isrc = snapshot_solution(net)
    cur = [net.gate(vg, gname) for gname in ("back", "partial", "total") for vg in (-10.0, 0.0, 10.0)]
    np.savez_compressed(os.path.join(HERE, name + "_answers.npz"), V0=V, isrc0=np.float64(isrc),
                        sweep_current=np.array(cur, dtype=np.float64), percolating=np.bool_(net.percolating),
                        onoffmap=np.int64(net.onoffmap), scaling=np.float64(net.scaling))
    print("%-28s dumped by the reference: %d sticks, %d junctions, isrc0=%r" % (name, len(net.sticks), len(net.intersects), isrc))


def main():
    if not refshim.available():
        sys.exit("reference not available; goldens can only be generated in the build container")
    dump_reference_system("refdump_s5_n300_seed1", n=300, scaling=5, seed=1, onoffmap=0)
    if "--dump-only" in sys.argv:
        return
    build("trivial", trivial=True)
    build("trivial_om1", trivial=True, onoffmap=1)
    build("s5_n300_seed1", n=300, scaling=5, seed=1, onoffmap=0)
    build("s5_n300_seed2_om1", n=300, scaling=5, seed=2, onoffmap=1)
    build("s5_n250_seed3_om2", n=250, scaling=5, seed=3, onoffmap=2)
    build("s5_n120_seed5_noperc", n=120, scaling=5, seed=5, onoffmap=0)
    build("s5_n300_seed6_fd", n=300, scaling=5, seed=6, onoffmap=0, element="FermiDiracTransistor")
    build("s5_n300_seed8_l1", n=300, scaling=5, seed=8, onoffmap=0, l=1.0)
    build("s20_n3600_seed7", n=3600, scaling=20, seed=7, onoffmap=1, sweep=False)
    build("s20_n6400_seed5", n=6400, scaling=20, seed=5, onoffmap=1, sweep=False)


if __name__ == "__main__":
    main()
