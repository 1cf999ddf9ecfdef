"""GPU: the two PCG implementations (multi-launch `flat`, persistent thread-block-cluster
`cluster`) against the refined oracle and against each other; Z-order row numbering."""
import os

import numpy as np
import pytest

from conftest import golden_sticks, load_golden
from oracle import cntnet_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture()
def solver(engine):
    old, oldp = engine.solver, engine.precond
    yield engine
    engine.solver, engine.precond = old, oldp
    os.environ.pop("CNT_CLUSTER_SIZE", None)
    os.environ.pop("CNT_CLUSTER_NOMB", None)
    os.environ.pop("CNT_AGG_HOST", None)


def _oracle_currents(g, onoffmap):
    s = golden_sticks(g)
    j = dict(stick1=g["stick1"], stick2=g["stick2"], x=g["jx"], y=g["jy"], kind2=g["kind2"], ntests=0)
    return O.measure_fullnet(s, onoffmap=onoffmap, junctions=j, refine=3)


@pytest.mark.parametrize("method", ["flat", "cluster"])
def test_methods_vs_oracle(solver, method):
    solver.solver = method
    names = ["trivial", "s5_n300_seed1", "s5_n120_seed5_noperc", "s5_n300_seed2_om1", "s20_n6400_seed5"]
    gs = [load_golden(n) for n in names]
    b = solver.batch_from_host([golden_sticks(g) for g in gs])
    res = solver.measure_fullnet(b, onoffmap=1)
    assert solver.last_solver == method
    assert np.all(res["status"] == 0)
    for i, g in enumerate(gs):
        m = _oracle_currents(g, 1)
        for k in ("ion", "ioff_back", "ioff_total", "ioff_partial"):
            if m["percolating"]:
                assert abs(res[k][i] - m[k]) <= 1e-9 * m[k], (method, g["name"], k, res[k][i], m[k])
            else:
                assert res[k][i] == 0.0


def test_flat_and_cluster_agree(solver):
    solver.precond = "jacobi"
    g = load_golden("s20_n6400_seed5")
    out = {}
    for method in ("flat", "cluster"):
        solver.solver = method
        b = solver.batch_from_host([golden_sticks(g)])
        out[method] = solver.measure_fullnet(b, onoffmap=1, want_fields=True)
    for q in range(4):
        Vf = out["flat"]["fields"]["V"][q].cpu().numpy()
        Vc = out["cluster"]["fields"]["V"][q].cpu().numpy()
        m = ~np.isnan(Vf)
        assert np.array_equal(np.isnan(Vf), np.isnan(Vc))
        assert np.max(np.abs(Vf[m] - Vc[m])) <= 1e-11 * 0.1
    # iteration counts agree up to the different (but fixed) summation trees
    assert np.all(np.abs(out["flat"]["iters"].astype(int) - out["cluster"]["iters"].astype(int)) <= 0.02 * out["flat"]["iters"] + 5)


@pytest.mark.parametrize("csize", [1, 2, 4, 8, 16])
def test_cluster_sizes(solver, csize):
    """the same systems on clusters of 1..16 CTAs (row chunks, DSMEM reductions, halo gathers)"""
    os.environ["CNT_CLUSTER_SIZE"] = str(csize)
    solver.solver = "cluster"
    names = ["s5_n300_seed1", "s20_n6400_seed5", "trivial"]
    gs = [load_golden(n) for n in names]
    b = solver.batch_from_host([golden_sticks(g) for g in gs])
    res = solver.measure_fullnet(b, onoffmap=1)
    assert np.all(res["status"] == 0)
    for i, g in enumerate(gs):
        m = _oracle_currents(g, 1)
        for k in ("ion", "ioff_back", "ioff_total", "ioff_partial"):
            assert abs(res[k][i] - m[k]) <= 1e-9 * m[k], (csize, g["name"], k)
    # deterministic: a second run gives the same bits
    b2 = solver.batch_from_host([golden_sticks(g) for g in gs])
    res2 = solver.measure_fullnet(b2, onoffmap=1)
    for k in ("ion", "ioff_back", "ioff_total", "ioff_partial"):
        assert np.array_equal(res[k], res2[k])
    assert np.array_equal(res["iters"], res2["iters"])


@pytest.mark.parametrize("chunk_C", [0, 2, 16])
def test_aggregate_tables_device_equals_host(solver, chunk_C):
    """segment / CTA tables of the two-level preconditioner: the device builder (flags + scans)
    reproduces the host reference builder entry by entry"""
    from networksim_cntfet_b200.elements import LinExpTransistor, conductance_table
    names = ["s5_n300_seed1", "s20_n6400_seed5", "s5_n120_seed5_noperc", "s20_n3600_seed7"]
    b = solver.prepare(solver.batch_from_host([golden_sticks(load_golden(n)) for n in names]))
    gates = [g for _, g in solver.fullnet_gates(b)]
    tabs = [conductance_table(LinExpTransistor, 1, g.levels) for g in gates]
    sysm = solver.assemble_values(b, gates, tabs)
    nets = [n for n in range(b.B) if b.h_roff[n + 1] > b.h_roff[n]]
    sys_net = [n for q in range(4) for n in nets]
    sys_slab = [q for q in range(4) for n in nets]
    out = {}
    for mode in ("device", "host"):
        os.environ.pop("CNT_AGG_HOST", None)
        if mode == "host":
            os.environ["CNT_AGG_HOST"] = "1"
        with solver._work():
            st, t = solver._aggregates(b, sysm, sys_net, sys_slab, chunk_C)
        solver.sync()
        out[mode] = (st.nagg, st.nseg, {k: v.cpu().numpy() for k, v in t.items()})
    os.environ.pop("CNT_AGG_HOST", None)
    (na, ns, d), (na2, ns2, h) = out["device"], out["host"]
    assert (na, ns) == (na2, ns2) and na > 0 and ns >= na
    nsys = len(sys_net)
    lens = dict(agg_of_row=None, mem=None, seg_start=ns, seg_count=ns, seg_agg=ns, seg_sys=ns, agg_seg0=na + 1,
                sys_seg=2 * nsys, row_q=None, row_einv=None, agg_einv=na)
    if chunk_C:
        nloc = int(h["cta_agg_off"][nsys * chunk_C])
        lens.update(cta_seg_off=nsys * chunk_C + 1, cta_seg=4 * ns, cta_agg_off=nsys * chunk_C + 1, cta_agg=nloc,
                    row_slot=None, seg_home=ns)
    for k, n in lens.items():
        a, c = (d[k], h[k]) if n is None else (d[k][:n], h[k][:n])
        assert np.array_equal(a, c), (chunk_C, k, np.flatnonzero(a != c)[:5])


def test_transaction_barriers_equal_cluster_barriers(solver):
    """the iteration synchronises through mbarrier transaction counts (st.async pushes); the
    reference path with cluster barriers does the same arithmetic -> identical bits"""
    solver.solver = "cluster"
    names = ["s5_n300_seed1", "s20_n6400_seed5", "trivial", "s20_n3600_seed7"]
    gs = [load_golden(n) for n in names]
    out = {}
    for mode in ("mb", "barrier"):
        os.environ.pop("CNT_CLUSTER_NOMB", None)
        if mode == "barrier":
            os.environ["CNT_CLUSTER_NOMB"] = "1"
        for csize in (2, 16):
            os.environ["CNT_CLUSTER_SIZE"] = str(csize)
            b = solver.batch_from_host([golden_sticks(g) for g in gs])
            out[mode, csize] = solver.measure_fullnet(b, onoffmap=1)
            assert solver.last_solver == "cluster"
    for csize in (2, 16):
        for k in ("ion", "ioff_back", "ioff_total", "ioff_partial", "iters"):
            assert np.array_equal(out["mb", csize][k], out["barrier", csize][k]), (csize, k)


def test_many_systems_work_queue(solver):
    """more systems than resident clusters: every cluster loops over the work queue"""
    solver.solver = "cluster"
    ns = [300 + 7 * k for k in range(60)]
    b = solver.generate(ns, 5.0, seed=5)
    res = solver.measure_fullnet(b, onoffmap=1)
    solver.solver = "flat"
    b2 = solver.generate(ns, 5.0, seed=5)
    ref = solver.measure_fullnet(b2, onoffmap=1)
    assert np.array_equal(res["percolating"], ref["percolating"]) and res["percolating"].sum() > 10
    for k in ("ion", "ioff_back", "ioff_total", "ioff_partial"):
        assert np.allclose(res[k], ref[k], rtol=1e-10, atol=0)
    assert np.all(res["status"] == 0) and np.all(ref["status"] == 0)


def test_maxit_and_status(solver):
    g = load_golden("s20_n6400_seed5")
    for method in ("flat", "cluster"):
        solver.solver = method
        b = solver.batch_from_host([golden_sticks(g)])
        res = solver.measure_fullnet(b, onoffmap=1, maxit=50, check_every=10)
        assert np.all(res["status"] == 1) and np.all(res["iters"] == 50), (method, res["status"], res["iters"])
        assert np.all(res["relres"] > 1e-15)


def test_z_order_locality(solver):
    """cnt_assemble_reorder: neighbouring sticks get neighbouring rows"""
    b = solver.generate([100000], 100.0, seed=1)
    solver.prepare(b)
    rp = b.rowptr.cpu().numpy()
    col = b.col.cpu().numpy()[:b.NNZ]
    rows = np.repeat(np.arange(b.R), np.diff(rp))
    dist = np.abs(col - rows)
    assert np.median(dist) < b.R / 100            # random order would give ~R/3
    chunk = (b.R + 15) // 16
    local = (col // chunk) == (rows // chunk)
    assert local.mean() > 0.85                    # gathers served from the CTA's own chunk (16-CTA cluster)
    solver.spatial_order = False
    try:
        b2 = solver.generate([100000], 100.0, seed=1)
        solver.prepare(b2)
        rp2 = b2.rowptr.cpu().numpy()
        col2 = b2.col.cpu().numpy()[:b2.NNZ]
        rows2 = np.repeat(np.arange(b2.R), np.diff(rp2))
        assert np.median(np.abs(col2 - rows2)) > 20 * np.median(dist)
        assert np.array_equal(b2.row_stick.cpu().numpy()[:b2.R], np.sort(b.row_stick.cpu().numpy()[:b.R]))
    finally:
        solver.spatial_order = True


def test_cluster_v1_uncompressed_path(solver):
    """the uncompressed cluster kernel (values streamed from L2) stays available: it serves
    caller-supplied conductances (cnet.ConductionNetwork built from a networkx graph)"""
    solver.solver = "cluster"
    solver.cluster_compressed = False
    try:
        g = load_golden("s20_n6400_seed5")
        b = solver.batch_from_host([golden_sticks(g)])
        res = solver.measure_fullnet(b, onoffmap=1)
        m = _oracle_currents(g, 1)
        for k in ("ion", "ioff_back", "ioff_total", "ioff_partial"):
            assert abs(res[k][0] - m[k]) <= 1e-9 * m[k]
    finally:
        solver.cluster_compressed = True


def test_cluster_overflow_falls_back_to_flat(solver):
    """a dense network (38 junctions per stick) does not fit the CTA's shared-memory slice:
    the cluster kernel reports status 3 and the engine re-solves with the multi-launch kernels"""
    solver.solver = "cluster"
    os.environ["CNT_CLUSTER_SIZE"] = "2"      # holds the rows but not the entries
    b = solver.generate([6000], 10.0, seed=9, l=1.0)
    res = solver.measure_fullnet(b, onoffmap=1)
    assert solver.last_solver == "cluster+flat"
    assert res["percolating"][0] and np.all(res["status"] == 0)
    s = solver.sticks_to_host(b)[0]
    m = O.measure_fullnet(s, onoffmap=1, refine=3)
    for k in ("ion", "ioff_back", "ioff_total", "ioff_partial"):
        assert abs(res[k][0] - m[k]) <= 1e-9 * m[k]


@pytest.mark.parametrize("method", ["flat", "cluster"])
@pytest.mark.parametrize("n,scaling", [(6400, 20.0), (100000, 100.0)])
def test_two_level_preconditioner(solver, n, scaling, method):
    """M^-1 = D^-1 + Z diag(Z^T A Z)^-1 Z^T on strong-link aggregates: same answers as plain
    Jacobi-PCG and as the refined oracle, in far fewer iterations; bit-reproducible"""
    solver.solver = method
    out = {}
    for precond in ("jacobi", "twolevel"):
        if precond == "jacobi" and n > 50000:
            continue                                  # 340k iterations: covered by the cluster tests
        solver.precond = precond
        b = solver.generate([n], scaling, seed=3)
        out[precond] = solver.measure_fullnet(b, onoffmap=1)
        assert np.all(out[precond]["status"] == 0)
    s = solver.sticks_to_host(b)[0]
    m = O.measure_fullnet(s, onoffmap=1, refine=3)
    two = out["twolevel"]
    assert two["percolating"][0] == m["percolating"]
    for k in ("ion", "ioff_back", "ioff_total", "ioff_partial"):
        assert abs(two[k][0] - m[k]) <= 1e-9 * m[k], (k, two[k][0], m[k], two["iters"][:, 0])
    assert solver.last_aggregates[0] > 0
    if "jacobi" in out:
        jac = out["jacobi"]
        for k in ("ion", "ioff_back", "ioff_total", "ioff_partial"):
            assert abs(two[k][0] - jac[k][0]) <= 1e-9 * jac[k][0]
        assert two["iters"][1, 0] * 4 < jac["iters"][1, 0]      # back gate: >4x fewer iterations
    else:
        assert two["iters"].max() < 12000                       # plain Jacobi: ~200k at Vg=+10
    again = solver.measure_fullnet(solver.generate([n], scaling, seed=3), onoffmap=1)
    for k in ("ion", "ioff_back", "ioff_total", "ioff_partial"):
        assert again[k][0] == two[k][0]
    assert np.array_equal(again["iters"], two["iters"])


def test_cluster_without_spatial_order(solver):
    """rows in stick (length) order instead of Z-order: the members of small aggregates land in different CTA
    chunks, so the segment tables of the cluster kernel need up to one segment per aggregated row -- sized and
    checked through seg_capacity (cnt_pcg_aggregates); same currents as with the spatial order"""
    solver.solver = "cluster"
    res = {}
    for order in (True, False):
        solver.spatial_order = order
        try:
            b = solver.generate([100000], 100.0, seed=5)
            res[order] = solver.measure_fullnet(b, onoffmap=1)
        finally:
            solver.spatial_order = True
        assert np.all(res[order]["status"] == 0)
        nagg, nseg = solver.last_aggregates
        assert nagg > 0 and nseg >= nagg
    for k in ("ion", "ioff_back", "ioff_total", "ioff_partial"):
        assert abs(res[False][k][0] - res[True][k][0]) <= 1e-9 * res[True][k][0]
