"""GPU: parity of the CUDA path (through the C-ABI) against the golden vectors produced by
the unmodified reference and against the numpy oracle on the same seeded inputs.

Bars: junction sets, junction coordinates, kind codes, cluster labels and cluster statistics
BIT-EXACT; voltages within 1e-9 * Vds (norm-wise), edge currents within 1e-9 relative plus the
voltage tolerance times the conductance, source current within 1e-9 relative of the refined
oracle (the raw reference itself is only ~2e-9 accurate at Vg=+10).
"""
import numpy as np
import pytest

from conftest import FULLNET_GATES, golden_element, golden_names, golden_sticks, load_golden
from oracle import cntnet_oracle as O
from reductions_ref import two_core, series_set  # noqa: E402

pytestmark = pytest.mark.gpu
NAMES = golden_names()
VDS = 0.1


def _element(name):
    from networksim_cntfet_b200 import elements
    return elements.FermiDiracTransistor if name == "fermidirac" else elements.LinExpTransistor


def _junction_slices(b):
    joff = b.joff.cpu().numpy()
    arr = {k: getattr(b, k).cpu().numpy() for k in ("stick1", "stick2", "jx", "jy", "kind2")}
    return [{k: v[joff[i]:joff[i + 1]] for k, v in arr.items()} for i in range(b.B)]


def _check_junctions(j, g):
    assert len(j["stick1"]) == len(g["stick1"]), (len(j["stick1"]), len(g["stick1"]))
    assert np.array_equal(j["stick1"], g["stick1"])
    assert np.array_equal(j["stick2"], g["stick2"])
    assert np.array_equal(j["jx"], g["jx"])      # bit-exact FP64 coordinates
    assert np.array_equal(j["jy"], g["jy"])
    assert np.array_equal(j["kind2"], g["kind2"])


@pytest.mark.parametrize("name", NAMES)
def test_junctions_single(engine, name):
    g = load_golden(name)
    b = engine.find_junctions(engine.batch_from_host([golden_sticks(g)]))
    _check_junctions(_junction_slices(b)[0], g)
    ref_tests = O.find_junctions(golden_sticks(g))["ntests"]
    assert int(b.ntests.cpu()[0]) == ref_tests          # same candidate set as the KD-tree ball


def test_junctions_and_clusters_ragged_batch(engine):
    """all goldens in ONE batch (ragged sizes 6 ... 6402 sticks, percolating or not)"""
    gs = [load_golden(n) for n in NAMES]
    b = engine.batch_from_host([golden_sticks(g) for g in gs])
    engine.find_junctions(b)
    engine.label_clusters(b)
    js = _junction_slices(b)
    label = b.label.cpu().numpy()
    stats = b.stats.cpu().numpy().reshape(b.B, -1)
    from networksim_cntfet_b200 import _lib
    for i, g in enumerate(gs):
        _check_junctions(js[i], g)
        lab = label[b.h_soff[i]:b.h_soff[i + 1]]
        assert np.array_equal(lab, g["label"]), g["name"]
        assert bool(stats[i, _lib.STAT_PERCOLATING]) == bool(g["percolating"])
        assert stats[i, _lib.STAT_NCLUST_REF] == int(g["nclust_ref"])
        assert stats[i, _lib.STAT_MAXCLUST_REF] == int(g["maxclust_ref"])
        sizes = np.bincount(g["label"])
        assert stats[i, _lib.STAT_NCLUST_TRUE] == int((sizes > 0).sum())
        assert stats[i, _lib.STAT_MAXCLUST_TRUE] == int(sizes.max())
        if g["percolating"]:
            assert stats[i, _lib.STAT_NCOMP] == int((g["label"] == 0).sum())
            assert stats[i, _lib.STAT_ECOMP] == int((g["label"][g["stick1"]] == 0).sum())


def test_tile_buffer_overflow_rows_are_recomputed(engine):
    """dense networks (constant stick length 1 um on 10x10 um: ~38 junctions per stick) overflow the tile kernel's
    hit buffer and the temporary table: exactly the rows of the affected home-stick batches are recomputed by
    k_cells_fill; a sparse network in the same batch keeps the fast path.  Sets, coordinates, ntests vs the oracle."""
    b = engine.generate([6000, 300, 3000], [10.0, 5.0, 8.0], seed=9, l=1.0)
    engine.find_junctions(b)
    assert b.junction_rows_recomputed
    js = _junction_slices(b)
    nt = b.ntests.cpu().numpy()
    for i, s in enumerate(engine.sticks_to_host(b)):
        j = O.find_junctions(s)
        assert len(j["stick1"]) > 0
        _check_junctions(js[i], dict(stick1=j["stick1"], stick2=j["stick2"], jx=j["x"], jy=j["y"], kind2=j["kind2"]))
        assert int(nt[i]) == j["ntests"]


def test_scan(engine):
    import torch
    from networksim_cntfet_b200 import _lib
    from networksim_cntfet_b200.engine import _p
    lib = _lib.load()
    rng = np.random.default_rng(0)
    for n in (1, 5, 4096, 4097, 100003, 3_000_001):
        a = rng.integers(0, 7, size=n).astype(np.int32)
        with torch.cuda.stream(engine.stream):
            t = torch.as_tensor(a).to(engine.device)
            out = torch.empty(n + 1, dtype=torch.int32, device=engine.device)
            ws = engine._ws(lib.cnt_scan_workspace_bytes(n))
            _lib.check(lib.cnt_exclusive_scan_i32(_p(t), _p(out), n, 1, _p(ws), ws.numel(), engine.stream_ptr))
        engine.sync()
        ref = np.concatenate([[0], np.cumsum(a)])
        assert np.array_equal(out.cpu().numpy(), ref)


def _oracle_fields(g, tag, gate, vg, refine=3):
    el, om = golden_element(g), int(g["onoffmap"])
    vgj = O.gate_voltages(g["jx"], g["jy"], gate, vg)
    cond = O.conductance(g["kind2"], vgj, om, el)
    r = O.solve_network(len(g["xc"]), g["stick1"], g["stick2"], cond, g["label"], refine=refine)
    return cond, r


PERC = [n for n in NAMES if "noperc" not in n and "n3600" not in n]


@pytest.mark.parametrize("name", PERC)
def test_assembly_matches_oracle(engine, name):
    """CSR of L_UU (pattern, values, diagonal, rhs) equals the oracle's matrix entry by entry."""
    from networksim_cntfet_b200.engine import GateState
    from networksim_cntfet_b200.elements import conductance_table
    g = load_golden(name)
    engine.prune = engine.series = False           # the reference's matrix: every stick of the component is a node
    try:
        b = engine.prepare(engine.batch_from_host([golden_sticks(g)]))
    finally:
        engine.prune = engine.series = "auto"
    gs = GateState(engine, b).set_global(10)
    tab = conductance_table(_element(golden_element(g)), int(g["onoffmap"]), gs.levels)
    sysm = engine.assemble_values(b, [gs], [tab])
    engine.sync()
    cond, r = _oracle_fields(g, "_back", "back", 10.0, refine=0)
    assert np.array_equal(sysm["g"][0].cpu().numpy()[:b.J], cond)       # bit-identical conductances
    R = r["A"].shape[0]
    assert b.R == R
    # the GPU numbers its unknowns in Z-order of the stick centres (cnt_assemble_reorder);
    # row_stick maps them back to the reference's ascending node order
    row_stick = b.row_stick.cpu().numpy()[:R]
    assert np.array_equal(np.sort(row_stick), r["nodes_u"])
    perm = np.searchsorted(r["nodes_u"], row_stick)
    A = r["A"].tocsr()[perm][:, perm].tocsr()
    A.sort_indices()
    r = dict(r, rhs=r["rhs"][perm], nodes_u=r["nodes_u"][perm])
    rowptr = b.rowptr.cpu().numpy()
    col = b.col.cpu().numpy()[:b.NNZ]
    val = sysm["val"][0].cpu().numpy()[:b.NNZ]
    diag = sysm["diag"][0].cpu().numpy()[:R]
    rhs = sysm["rhs"][0].cpu().numpy()[:R]
    import scipy.sparse as sp
    M = sp.csr_matrix((val, col, rowptr), shape=(R, R)) + sp.diags(diag)
    D = (M - A).tocoo()
    assert np.all(np.abs(D.data) <= 1e-15 * np.abs(A.diagonal()).max())
    offd = sp.csr_matrix((val, col, rowptr), shape=(R, R))
    offd.sort_indices()
    Aoff = (A - sp.diags(A.diagonal())).tocsr()
    Aoff.eliminate_zeros()
    Aoff.sort_indices()
    assert np.array_equal(offd.indptr, Aoff.indptr) and np.array_equal(offd.indices, Aoff.indices)
    assert np.array_equal(offd.data, Aoff.data)                           # off-diagonals bit-exact
    assert np.array_equal(rhs, r["rhs"])
    assert np.array_equal(b.row_stick.cpu().numpy()[:R], r["nodes_u"])


@pytest.mark.parametrize("name", PERC)
def test_pruned_assembly(engine, name):
    """cnt_prune_leaves: the unknowns are the 2-core of the percolating component, the matrix is
    the oracle's matrix restricted to them with the dangling junctions' conductance taken off the
    diagonal, and pruned sticks are attached to the stick they hang on."""
    import scipy.sparse as sp
    from networksim_cntfet_b200.engine import GateState
    from networksim_cntfet_b200.elements import conductance_table
    g = load_golden(name)
    engine.prune, engine.series = True, False
    try:
        b = engine.prepare(engine.batch_from_host([golden_sticks(g)]))
    finally:
        engine.prune = engine.series = "auto"
    gs = GateState(engine, b).set_global(10)
    tab = conductance_table(_element(golden_element(g)), int(g["onoffmap"]), gs.levels)
    sysm = engine.assemble_values(b, [gs], [tab])
    engine.sync()
    cond, r = _oracle_fields(g, "_back", "back", 10.0, refine=0)
    n = len(g["xc"])
    comp = g["label"] == 0
    alive, attach = two_core(n, g["stick1"].astype(np.int64), g["stick2"].astype(np.int64), comp)
    assert np.array_equal(b.alive.cpu().numpy().astype(bool), alive)
    assert np.array_equal(b.attach.cpu().numpy(), attach)
    keep = alive[r["nodes_u"]]
    R = int(keep.sum())
    assert b.R == R and R <= r["A"].shape[0]
    row_stick = b.row_stick.cpu().numpy()[:R]
    nodes = r["nodes_u"][keep]
    assert np.array_equal(np.sort(row_stick), nodes)
    perm = np.searchsorted(nodes, row_stick)
    A = r["A"].tocsr()[keep][:, keep].tocsr()[perm][:, perm].tocsr()
    offd = sp.csr_matrix((sysm["val"][0].cpu().numpy()[:b.NNZ], b.col.cpu().numpy()[:b.NNZ], b.rowptr.cpu().numpy()),
                         shape=(R, R))
    offd.sort_indices()
    Aoff = (A - sp.diags(A.diagonal())).tocsr()
    Aoff.eliminate_zeros()
    Aoff.sort_indices()
    assert np.array_equal(offd.indptr, Aoff.indptr) and np.array_equal(offd.indices, Aoff.indices)
    assert np.array_equal(offd.data, Aoff.data)
    # diagonal: the full one minus the conductance of the junctions to pruned sticks
    s1, s2 = g["stick1"].astype(np.int64), g["stick2"].astype(np.int64)
    dang = comp[s1] & comp[s2] & (alive[s1] != alive[s2])
    off = np.zeros(n)
    np.add.at(off, np.where(alive[s1[dang]], s1[dang], s2[dang]), cond[dang])
    want = A.diagonal() - off[row_stick]
    diag = sysm["diag"][0].cpu().numpy()[:R]
    assert np.all(np.abs(diag - want) <= 4e-16 * np.abs(A.diagonal()).max())
    assert np.array_equal(sysm["rhs"][0].cpu().numpy()[:R], r["rhs"][keep][perm])


@pytest.mark.parametrize("name", PERC)
def test_series_assembly(engine, name):
    """cnt_series_select: an independent set of degree-2 sticks of the 2-core leaves the system; the
    matrix is the Schur complement of the pruned matrix over them (one virtual junction of
    conductance g1 g2 / (g1 + g2) per eliminated stick)."""
    import scipy.sparse as sp
    from networksim_cntfet_b200.engine import GateState
    from networksim_cntfet_b200.elements import conductance_table
    g = load_golden(name)
    engine.prune = engine.series = True
    try:
        b = engine.prepare(engine.batch_from_host([golden_sticks(g)]))
    finally:
        engine.prune = engine.series = "auto"
    gs = GateState(engine, b).set_global(10)
    tab = conductance_table(_element(golden_element(g)), int(g["onoffmap"]), gs.levels)
    sysm = engine.assemble_values(b, [gs], [tab])
    engine.sync()
    cond, r = _oracle_fields(g, "_back", "back", 10.0, refine=0)
    n = len(g["xc"])
    s1, s2 = g["stick1"].astype(np.int64), g["stick2"].astype(np.int64)
    alive, _ = two_core(n, s1, s2, g["label"] == 0)
    want = series_set(n, s1, s2, alive)
    assert b.NV == len(want)
    if b.NV == 0:
        assert b.elim_vid is None
        return
    vid = b.elim_vid.cpu().numpy()[:n]
    assert np.array_equal(np.nonzero(vid >= 0)[0], np.array(sorted(want)))
    assert np.array_equal(vid[vid >= 0], np.arange(b.NV))                # numbered in stick order
    va, vc, ve1, ve2 = (t.cpu().numpy()[:b.NV] for t in (b.va, b.vc, b.ve1, b.ve2))
    for s, (n0, n1, e0, e1) in want.items():
        assert (va[vid[s]], vc[vid[s]], ve1[vid[s]], ve2[vid[s]]) == (n0, n1, e0, e1)
    assert b.voff.cpu().numpy().tolist() == [0, b.NV]
    elim = vid >= 0
    keep_a = alive[r["nodes_u"]]
    A = r["A"].tocsr()[keep_a][:, keep_a].tocsr()                            # pruned matrix, ascending sticks
    nodes = r["nodes_u"][keep_a]
    dang = (g["label"] == 0)[s1] & (g["label"] == 0)[s2] & (alive[s1] != alive[s2])
    off = np.zeros(n)
    np.add.at(off, np.where(alive[s1[dang]], s1[dang], s2[dang]), cond[dang])
    A = (A - sp.diags(off[nodes])).tocsr()
    ke, kk = elim[nodes], ~elim[nodes]
    Aee = A[ke][:, ke]
    assert (Aee - sp.diags(Aee.diagonal())).nnz == 0 or np.all((Aee - sp.diags(Aee.diagonal())).data == 0)   # independent
    S = (A[kk][:, kk] - A[kk][:, ke] @ sp.diags(1.0 / Aee.diagonal()) @ A[ke][:, kk]).tocsr()
    R = int(kk.sum())
    assert b.R == R
    row_stick = b.row_stick.cpu().numpy()[:R]
    assert np.array_equal(np.sort(row_stick), nodes[kk])
    perm = np.searchsorted(nodes[kk], row_stick)
    S = S[perm][:, perm].tocsr()
    M = sp.csr_matrix((sysm["val"][0].cpu().numpy()[:b.NNZ], b.col.cpu().numpy()[:b.NNZ], b.rowptr.cpu().numpy()),
                      shape=(R, R)) + sp.diags(sysm["diag"][0].cpu().numpy()[:R])
    D = (M - S).tocoo()
    assert np.all(np.abs(D.data) <= 4e-16 * np.abs(A.diagonal()).max())
    assert np.array_equal(sysm["rhs"][0].cpu().numpy()[:R], r["rhs"][keep_a][kk][perm])
    # class-compressed values of the cluster kernel reproduce val exactly
    if "cls" in sysm:
        cls = sysm["cls"][0].cpu().numpy()[:b.NNZ]
        gneg = sysm["gneg"][0].cpu().numpy() if hasattr(sysm["gneg"], "cpu") else np.asarray(sysm["gneg"])[0]
        assert np.array_equal(gneg[cls], sysm["val"][0].cpu().numpy()[:b.NNZ])


@pytest.mark.parametrize("name", PERC)
def test_fullnet_solve(engine, name):
    """measure_fullnet gate sequence: V, |I|, I_source against refined oracle and raw reference."""
    g = load_golden(name)
    b = engine.batch_from_host([golden_sticks(g)])
    res = engine.measure_fullnet(b, onoffmap=int(g["onoffmap"]), element=_element(golden_element(g)),
                                 want_fields=True)
    assert res["percolating"][0]
    assert np.all(res["status"] == 0), res["status"]
    f = res["fields"]
    for q, (tag, gate, vg) in enumerate(FULLNET_GATES):
        cond, r = _oracle_fields(g, tag, gate, vg)
        V = f["V"][q].cpu().numpy()[:b.S]
        I = f["I"][q].cpu().numpy()[:b.J]
        assert np.array_equal(np.isnan(V), np.isnan(r["V"]))
        m = ~np.isnan(r["V"])
        assert np.max(np.abs(V[m] - r["V"][m])) <= 1e-9 * VDS, (tag, np.max(np.abs(V[m] - r["V"][m])))
        me = ~np.isnan(r["I"])
        assert np.array_equal(np.isnan(I), ~me)
        assert np.all(np.abs(I[me] - r["I"][me]) <= 1e-9 * np.abs(r["I"][me]) + cond[me] * 2e-10 * VDS)
        isrc = res["isrc_all"][q, 0]
        # power form vs refined oracle: 1e-9 relative (north star); drain-side sum agrees too
        assert abs(isrc[2] - r["I_source"]) <= 1e-9 * r["I_source"], (tag, isrc, r["I_source"])
        assert abs(isrc[1] - r["I_source"]) <= 1e-7 * r["I_source"]
        assert abs(isrc[0] - r["I_source"]) <= 1e-5 * r["I_source"]
        # raw reference (its own solve error dominates at Vg=+10)
        ref = float(g["isrc0" if tag == "0" else "isrc" + tag])
        assert abs(isrc[2] - ref) <= (5e-9 if gate == "back" else 1e-9) * ref
        Vref = g["V0" if tag == "0" else "V" + tag]
        assert np.max(np.abs(V[m] - Vref[m])) <= 1e-9 * VDS


def test_fullnet_batch_matches_single(engine):
    """A ragged batch gives the same per-network rows as one-network batches (deterministic,
    batch-composition independent reductions), and non-percolating networks report zeros."""
    names = ["s5_n300_seed1", "s5_n120_seed5_noperc", "trivial", "s5_n300_seed2_om1", "s20_n3600_seed7"]
    gs = [load_golden(n) for n in names]
    b = engine.batch_from_host([golden_sticks(g) for g in gs])
    res = engine.measure_fullnet(b, onoffmap=1)
    for i, g in enumerate(gs):
        one = engine.measure_fullnet(engine.batch_from_host([golden_sticks(g)]), onoffmap=1)
        for k in ("ion", "ioff_back", "ioff_total", "ioff_partial"):
            assert res[k][i] == one[k][0], (g["name"], k)            # bit-identical
        assert res["percolating"][i] == bool(g["percolating"])
        if not g["percolating"]:
            assert res["ion"][i] == 0.0 and res["ioff_back"][i] == 0.0
        assert res["nclust"][i] == int(g["nclust_ref"]) and res["maxclust"][i] == int(g["maxclust_ref"])


def test_edge_cases(engine):
    """empty-ish networks: electrodes only; a network without any junction; parallel sticks."""
    e = dict(xc=np.array([0.01, 0.99]), yc=np.array([0.5, 0.5]), angle=np.full(2, np.pi / 2 - 1e-6),
             length=np.array([100.0, 100.0]), kind=np.zeros(2, np.uint8))
    e = O.sticks_with_ends(e)
    e["orig"] = np.array([0, 1])
    lone = dict(xc=np.array([0.01, 0.99, 0.5]), yc=np.array([0.5, 0.5, 0.5]),
                angle=np.array([np.pi / 2 - 1e-6, np.pi / 2 - 1e-6, 0.3]), length=np.array([100.0, 100.0, 0.1]),
                kind=np.array([0, 0, 2], np.uint8))
    lone = O.sticks_with_ends(lone)
    lone["orig"] = np.array([0, 2, 1])
    vert = dict(xc=np.array([0.01, 0.99, 0.5, 0.5]), yc=np.array([0.5, 0.5, 0.5, 0.5]),
                angle=np.array([np.pi / 2 - 1e-6, np.pi / 2 - 1e-6, 0.0, 0.0]),
                length=np.array([100.0, 100.0, 0.99, 0.5]), kind=np.array([0, 0, 1, 2], np.uint8))
    vert = O.sticks_with_ends(vert)   # two collinear horizontal sticks: equal slopes -> no junction between them
    vert["orig"] = np.array([0, 3, 1, 2])
    b = engine.batch_from_host([e, lone, vert])
    res = engine.measure_fullnet(b, onoffmap=0)
    for i, s in enumerate((e, lone, vert)):
        j = O.find_junctions(s, pruner="brute")
        c = O.label_clusters(len(s["xc"]), j["stick1"], j["stick2"], s["orig"])
        assert res["njunctions"][i] == len(j["stick1"])
        assert res["percolating"][i] == c["percolating"]
        assert res["nclust"][i] == c["nclust_ref"] and res["maxclust"][i] == c["maxclust_ref"]
    # `vert`: the long horizontal stick touches both electrodes -> percolates through one node
    m = O.measure_fullnet(vert, onoffmap=0, refine=2)
    assert res["percolating"][2] and abs(res["ion"][2] - m["ion"]) <= 1e-12 * m["ion"]


def test_unsorted_table_is_rejected(engine):
    from networksim_cntfet_b200._lib import CntError
    g = load_golden("s5_n300_seed1")
    s = golden_sticks(g)
    s = {k: v[::-1].copy() for k, v in s.items()}
    with pytest.raises(CntError):
        engine.find_junctions(engine.batch_from_host([s]))


@pytest.mark.parametrize("n,scaling,seed", [(2500, 20, 11), (30000, 60, 12)])
def test_seeded_vs_oracle(engine, n, scaling, seed):
    """same seeded inputs, sizes the oracle finishes in seconds: full pipeline vs oracle"""
    s = O.sticks_with_ends(O.canonical_sort(O.make_sticks_mt19937(n, "exp", 0.135, scaling, seed)))
    b = engine.batch_from_host([s])
    res = engine.measure_fullnet(b, onoffmap=1)
    j = O.find_junctions(s)
    js = _junction_slices(b)[0]
    _check_junctions(js, dict(stick1=j["stick1"], stick2=j["stick2"], jx=j["x"], jy=j["y"], kind2=j["kind2"]))
    assert int(res["ntests"][0]) == j["ntests"]
    c = O.label_clusters(n + 2, j["stick1"], j["stick2"], s["orig"])
    assert np.array_equal(b.label.cpu().numpy()[:n + 2], c["label"])
    m = O.measure_fullnet(s, onoffmap=1, junctions=j, refine=3)
    assert res["percolating"][0] == m["percolating"]
    assert res["nclust"][0] == m["nclust_ref"] and res["maxclust"][0] == m["maxclust_ref"]
    for k in ("ion", "ioff_back", "ioff_total", "ioff_partial"):
        if m["percolating"]:
            assert abs(res[k][0] - m[k]) <= 1e-9 * m[k], (k, res[k][0], m[k])
        else:
            assert res[k][0] == 0.0
