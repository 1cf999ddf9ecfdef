"""Direct solve of the junction-graph systems by parallel star-mesh elimination.

The reduced MNA system of a network (cnet.py:104-149) is a weighted graph Laplacian plus positive
couplings to the two electrodes.  Eliminating an unknown v with neighbours u_i (conductances g_i)
and electrode coupling c_v is the star-mesh transform:

    d_v = c_v + sum_i g_i,   new junction u_i - u_j: g_i g_j / d_v,
    c_ui += g_i c_v / d_v,   b_ui += g_i b_v / d_v,        afterwards x_v = (b_v + sum_i g_i x_ui) / d_v

i.e. sparse Cholesky written on the conductances.  Nothing is ever subtracted -- every quantity is
a sum of positive terms -- so the elimination is componentwise accurate (the GTH argument) and does
not suffer from the conditioning of the nearly floating gated networks that limits both PCG and
SuperLU.  Close to the percolation threshold, where the reference's ensembles
live, the current-carrying backbone is almost tree-like: a minimum-degree elimination of a
100x100 um system (46k unknowns) creates only ~1.4x fill and ~1e6 multiply-adds, against 1.7e8 row
updates of the 3.7k PCG iterations.

Parallel form: in every round an independent set of low-degree unknowns (lowest (degree, hash) among
their candidate neighbours) is eliminated at once for all networks of the batch.  The *symbolic*
phase below depends on the topology only (shared by all gate configurations): it works on sorted
edge keys with torch integer ops on the device and emits, per round, the index lists of three small
numeric kernels (pivot sums, new edge values, node updates) and of the back-substitution, each a
fixed-order segmented sum -> deterministic.  The numeric phase runs in CUDA (csrc/starmesh.cu) for
all gate configurations together.
"""
import ctypes as C

import torch


class Round(object):
    """index lists of one elimination round (all int64 tensors on the plan's device)"""
    __slots__ = ("E", "nptr", "nbr", "eslot", "nE_old", "nE_new", "copy_src", "cptr", "cm", "ca", "cb",
                 "tn", "tptr", "tm", "tslot", "gbase")


class Plan(object):
    """symbolic factorisation of a batch pattern"""

    def __init__(self):
        self.rounds = []
        self.R = 0
        self.e0_ptr = self.e0_src = None        # round-0 edges: sums of the CSR entries of a pair
        self.nE0 = 0
        self.nnzL = 0                           # sum of the degrees of the eliminated unknowns
        self.max_deg = 0


def _count(idx, n):
    """histogram of idx over [0, n) without the host synchronisation of torch.bincount"""
    return torch.zeros(int(n), dtype=torch.int64, device=idx.device).scatter_add_(0, idx, torch.ones_like(idx))


def _ragged_arange(counts):
    """[0..c0-1, 0..c1-1, ...] and the owner index of every element"""
    owner = torch.repeat_interleave(torch.arange(counts.numel(), device=counts.device), counts)
    start = torch.cumsum(counts, 0) - counts
    return torch.arange(owner.numel(), device=counts.device) - start[owner], owner


def symbolic(rowptr, col, R, max_fill=None, local_ids=None, net_ids=None, nnet=1, dense_max=0):
    """rowptr [R+1], col [NNZ] (global row indices of the batch; symmetric pattern, no diagonal).
    Every round eliminates the unknowns whose (degree, hash, index) is the smallest among their
    neighbours: an independent set, low degrees first (a parallel minimum-degree ordering; measured on
    a 100x100 um system: 85 rounds, 1.6 entries of L per entry of A, largest pivot degree 18 --
    the same fill as a sequential minimum-degree ordering).  Returns a Plan, or None when the number of
    star-mesh contributions exceeds max_fill times the number of edges (dense, mesh-like networks far
    above the percolation threshold: the caller then uses an iterative solver).
    local_ids [R]: index of every row inside its own network; the tie-break hash is taken from it, so a
    network is eliminated in the same order -- and gives the same bits -- whatever else is in the batch.
    net_ids [R], nnet, dense_max: as soon as no network has more than dense_max unknowns left, the rounds
    stop and the rest -- the dense core the rounds would take a handful at a time -- is described as one
    small dense system per network (plan.dense; csrc/starmesh.cu k_sm_dense)."""
    dev = rowptr.device
    i64 = torch.int64
    rowptr = rowptr.to(i64)
    col = col.to(i64)
    row = torch.repeat_interleave(torch.arange(R, device=dev, dtype=i64), rowptr.diff())
    upper = row < col
    nz = torch.nonzero(upper).flatten()
    key = row[nz] * R + col[nz]
    ukey, inv = torch.unique(key, return_inverse=True)
    order = torch.sort(inv, stable=True)[1]
    plan = Plan()
    plan.R = R
    plan.nE0 = int(ukey.numel())
    plan.e0_src = nz[order]
    plan.e0_ptr = torch.cat([torch.zeros(1, dtype=i64, device=dev), torch.bincount(inv, minlength=plan.nE0).cumsum(0)])
    curkey = ukey
    eu, ew = ukey // R, ukey % R
    deg = _count(eu, R) + _count(ew, R)                  # maintained incrementally below
    alive = torch.ones(R, dtype=torch.bool, device=dev)
    ids = torch.arange(R, device=dev, dtype=i64) if local_ids is None else local_ids.to(i64)
    h = (ids * 2654435761) % 1048573                     # fixed pseudo-random tie break
    gbase = 0
    n_alive = R
    npairs = 0
    maxdeg_t = torch.zeros((), dtype=i64, device=dev)
    plan.dense = None
    if net_ids is None:
        net_ids = torch.zeros(R, dtype=i64, device=dev)
    while n_alive > 0:
        if dense_max and n_alive <= dense_max * nnet:
            an = torch.nonzero(alive).flatten()
            nid = net_ids[an]
            cntb = torch.zeros(nnet, dtype=i64, device=dev).scatter_add_(0, nid, torch.ones_like(nid))
            if int(cntb.max()) <= dense_max:
                dptr = torch.cat([torch.zeros(1, dtype=i64, device=dev), cntb.cumsum(0)])
                loc = torch.full((R,), -1, dtype=i64, device=dev)
                loc[an] = torch.arange(an.numel(), device=dev, dtype=i64) - dptr[nid]
                doff = torch.cat([torch.zeros(1, dtype=i64, device=dev), (cntb * cntb).cumsum(0)])
                enet = net_ids[eu]
                plan.dense = dict(dptr=dptr, dnodes=an, doff=doff, epos=doff[enet] + loc[eu] * cntb[enet] + loc[ew],
                                  nE=int(eu.numel()), nmax=int(cntb.max()), gpool=max(int(doff[-1]), 1), nnet=int(nnet))
                break
        nE = int(eu.numel())
        pri = (deg * 1048573 + h) * R + ids
        loser = torch.where(pri[eu] > pri[ew], eu, ew)     # every edge of the list joins two live unknowns
        blocked = torch.zeros(R, dtype=torch.bool, device=dev)
        blocked[loser] = True
        sel = alive & ~blocked
        E = torch.nonzero(sel).flatten()
        mid = torch.full((R,), -1, dtype=i64, device=dev)
        mid[E] = torch.arange(E.numel(), device=dev, dtype=i64)
        su, sw = sel[eu], sel[ew]
        inc = su | sw
        slot = torch.nonzero(inc).flatten()
        v = torch.where(su[slot], eu[slot], ew[slot])
        o = torch.where(su[slot], ew[slot], eu[slot])
        srt = torch.argsort(v * R + o)
        v, o, slot = v[srt], o[srt], slot[srt]
        m = mid[v]
        cnt = _count(m, E.numel())
        nptr = torch.cat([torch.zeros(1, dtype=i64, device=dev), cnt.cumsum(0)])
        # all pairs (i < j) of every eliminated unknown's neighbours
        pos = torch.arange(v.numel(), device=dev, dtype=i64) - nptr[m]
        partners = cnt[m] - 1 - pos
        npairs += int(partners.sum())
        if max_fill is not None and npairs > max_fill * max(plan.nE0, 1) + 100000:
            return None
        off, ia = _ragged_arange(partners)
        ib = ia + off + 1
        pkey = o[ia] * R + o[ib]                          # neighbours ascending inside a node: o[ia] < o[ib]
        keep = ~inc
        kept = torch.nonzero(keep).flatten()
        okey = curkey[kept]                               # ascending (subset of the sorted key list)
        # merge the surviving edges with the NEW pairs instead of sorting everything again: the pairs are few.
        # One stable sort of the pair keys serves both the list of distinct pairs and the grouping of the
        # contributions by target edge
        spk, po = torch.sort(pkey, stable=True)
        if spk.numel():
            firstp = torch.ones(spk.numel(), dtype=torch.bool, device=dev)
            firstp[1:] = spk[1:] != spk[:-1]
            pk = spk[firstp]
        else:
            pk = spk
        if okey.numel() and pk.numel():
            at = torch.searchsorted(okey, pk).clamp(max=int(okey.numel()) - 1)
            pnew = pk[okey[at] != pk]
        else:
            pnew = pk
        opos = torch.arange(okey.numel(), device=dev, dtype=i64) + (torch.searchsorted(pnew, okey) if pnew.numel() else 0)
        ppos = torch.arange(pnew.numel(), device=dev, dtype=i64) + (torch.searchsorted(okey, pnew) if okey.numel() else 0)
        nE_new = int(okey.numel() + pnew.numel())
        newkey = torch.empty(nE_new, dtype=i64, device=dev)
        newkey[opos] = okey
        newkey[ppos] = pnew
        copy_src = torch.full((nE_new,), -1, dtype=i64, device=dev)
        copy_src[opos] = kept
        pslot = torch.searchsorted(newkey, spk) if spk.numel() else spk           # ascending, like spk
        cptr = torch.cat([torch.zeros(1, dtype=i64, device=dev),
                          _count(pslot, nE_new).cumsum(0)]) if nE_new else torch.zeros(1, dtype=i64, device=dev)
        # neighbours whose c, b change: one stable sort of the incident list by neighbour
        so, to = torch.sort(o, stable=True)
        if so.numel():
            firstt = torch.ones(so.numel(), dtype=torch.bool, device=dev)
            firstt[1:] = so[1:] != so[:-1]
            tstart = torch.nonzero(firstt).flatten()
            tn = so[tstart]
            tptr = torch.cat([tstart, torch.full((1,), so.numel(), dtype=i64, device=dev)])
        else:
            tn = so
            tptr = torch.zeros(1, dtype=i64, device=dev)
        r = Round()
        r.E, r.nptr, r.nbr, r.eslot = E, nptr, o, slot
        r.nE_old, r.nE_new = nE, nE_new
        r.copy_src, r.cptr = copy_src, cptr
        r.cm, r.ca, r.cb = m[ia][po], slot[ia][po], slot[ib][po]
        r.tn = tn
        r.tptr = tptr
        r.tm, r.tslot = m[to], slot[to]
        r.gbase = gbase
        gbase += int(v.numel())
        maxdeg_t = torch.maximum(maxdeg_t, cnt.max()) if cnt.numel() else maxdeg_t
        plan.rounds.append(r)
        alive[E] = False
        n_alive -= int(E.numel())
        # degrees: every neighbour loses its edge to the pivot and gains the NEW edges among the neighbours
        deg.scatter_add_(0, o, torch.full_like(o, -1))
        if pnew.numel():
            one = torch.ones_like(pnew)
            deg.scatter_add_(0, pnew // R, one)
            deg.scatter_add_(0, pnew % R, one)
        curkey = newkey
        eu, ew = newkey // R, newkey % R
    plan.nnzL = gbase
    plan.max_deg = int(maxdeg_t)
    return plan


# --------------------------------------------------------------------------- CUDA numeric phase
class SmRound(C.Structure):
    """mirror of cnt_sm_round (include/cntnet_b200.h)"""
    _fields_ = [("nelim", C.c_int64), ("nE_new", C.c_int64), ("ntarget", C.c_int64), ("ebase", C.c_int64),
                ("gbase", C.c_int64)] + [(n, C.c_void_p) for n in
                                         ("E", "nptr", "nbr", "eslot", "copy_src", "cptr", "cm", "ca", "cb", "tn", "tptr",
                                          "tm", "tslot")]


def finalize(plan, keep64=False):
    """int32 copies of the index lists in ONE device buffer + the host array of cnt_sm_round"""
    names = ("E", "nptr", "nbr", "eslot", "copy_src", "cptr", "cm", "ca", "cb", "tn", "tptr", "tm", "tslot")
    parts, offs, pos = [plan.e0_ptr, plan.e0_src], [], 0
    for r in plan.rounds:
        for n in names:
            parts.append(getattr(r, n))
    sizes = [int(t.numel()) for t in parts]
    pool = torch.cat([t.reshape(-1) for t in parts]).to(torch.int32)
    base = pool.data_ptr()
    starts = [0]
    for sz in sizes:
        starts.append(starts[-1] + sz)
    plan.pool = pool
    plan.e0_ptr_addr = base + 4 * starts[0]
    plan.e0_src_addr = base + 4 * starts[1]
    arr = (SmRound * max(len(plan.rounds), 1))()
    ebase, k = 0, 2
    for i, r in enumerate(plan.rounds):
        a = arr[i]
        a.nelim, a.nE_new, a.ntarget = int(r.E.numel()), int(r.nE_new), int(r.tn.numel())
        a.ebase, a.gbase = ebase, int(r.gbase)
        ebase += a.nelim
        for n in names:
            setattr(a, n, base + 4 * starts[k])
            k += 1
    plan.c_rounds = arr
    plan.nE_max = max([plan.nE0] + [int(r.nE_new) for r in plan.rounds] + [1])
    plan.max_deg = max(plan.max_deg, plan.dense["nmax"] if plan.dense else 0)
    # the symbolic lists are no longer needed on the device in 64-bit form
    plan.round_sizes = [int(r.E.numel()) for r in plan.rounds]
    plan.npairs = sum(int(r.cm.numel()) for r in plan.rounds)
    plan.nrounds = len(plan.rounds)
    if plan.dense is not None:
        d = plan.dense
        d["dptr32"] = d["dptr"].to(torch.int32)
        d["dnodes32"] = d["dnodes"].to(torch.int32)
        d["doff"] = d["doff"].contiguous()
        d["epos"] = d["epos"].contiguous()
    plan.pool_entries = int(pool.numel())
    plan.sum_nE = sum(int(r.nE_old) + int(r.nE_new) for r in plan.rounds)
    if not keep64:
        plan.rounds = None          # the 64-bit lists are not needed any more
        plan.e0_ptr = plan.e0_src = None
    return plan


def solve_cuda(plan, lib, stream_ptr, rowptr, val, gs, gd, rhs, leak=True):
    """Numeric phase on the GPU.  rowptr [R+1] int32; val [NS, >=NNZ] (CSR off-diagonals, -g);
    gs, gd [NS, R]: conductance of every unknown to the source / drain electrode (0 = none);
    rhs [NS, >=R].  Returns x [NS, R]."""
    from ._lib import check
    NS, R = gs.shape[0], plan.R
    dev = val.device
    f64 = torch.float64
    assert val.is_contiguous() and gs.is_contiguous() and gd.is_contiguous()
    g0 = torch.empty((NS, plan.nE_max), dtype=f64, device=dev)
    g1 = torch.empty((NS, plan.nE_max), dtype=f64, device=dev)
    c = torch.empty((NS, R), dtype=f64, device=dev)
    b = rhs[:, :R].contiguous().clone()
    dinv = torch.empty((NS, max(R, 1)), dtype=f64, device=dev)
    gE = torch.empty((NS, max(plan.nnzL, 1)), dtype=f64, device=dev)
    x = torch.zeros((NS, max(R, 1)), dtype=f64, device=dev)
    vp = C.c_void_p
    check(lib.cnt_sm_edges0(NS, plan.nE0, vp(plan.e0_ptr_addr), vp(plan.e0_src_addr), vp(val.data_ptr()), val.stride(0),
                            vp(g0.data_ptr()), plan.nE_max, stream_ptr))
    check(lib.cnt_sm_coupling(NS, R, vp(rowptr.data_ptr()), vp(val.data_ptr()), val.stride(0), vp(gs.data_ptr()),
                              vp(gd.data_ptr()), 1 if leak else 0, vp(c.data_ptr()), stream_ptr))
    check(lib.cnt_sm_forward(NS, plan.nrounds, plan.c_rounds, R, vp(g0.data_ptr()), vp(g1.data_ptr()), plan.nE_max,
                             vp(c.data_ptr()), vp(b.data_ptr()), vp(dinv.data_ptr()), dinv.stride(0), vp(gE.data_ptr()),
                             gE.stride(0), stream_ptr))
    if plan.dense is not None:
        d = plan.dense
        cur = g0 if plan.nrounds % 2 == 0 else g1
        G = torch.empty((NS, d["gpool"]), dtype=f64, device=dev)
        check(lib.cnt_sm_dense(NS, R, d["nnet"], vp(d["doff"].data_ptr()), vp(d["dptr32"].data_ptr()),
                               vp(d["dnodes32"].data_ptr()), d["nE"], vp(d["epos"].data_ptr()) if d["nE"] else None,
                               vp(cur.data_ptr()), plan.nE_max, vp(G.data_ptr()), d["gpool"], vp(c.data_ptr()),
                               vp(b.data_ptr()), vp(x.data_ptr()), d["nmax"], stream_ptr))
    check(lib.cnt_sm_backward(NS, plan.nrounds, plan.c_rounds, R, vp(b.data_ptr()), vp(dinv.data_ptr()), dinv.stride(0),
                              vp(gE.data_ptr()), gE.stride(0), vp(x.data_ptr()), stream_ptr))
    return x


def algorithmic_bytes(plan, NS):
    """bytes the direct stage has to move at least once: symbolic -- the sorted edge keys of every round
    read and written once (8 B each) and the index lists written once (4 B per entry); numeric -- the
    index lists read once, and per configuration the edge values of every round read and written
    once, the factor entries (gE) written and read, the three operands of every star-mesh contribution
    and 8 vectors of R doubles"""
    sym = 16 * plan.sum_nE + 4 * plan.pool_entries
    num = 4 * plan.pool_entries + 8 * NS * (plan.sum_nE + 3 * plan.nnzL + 3 * plan.npairs + 8 * plan.R)
    if plan.dense is not None:        # dense tail: every matrix written (zero + fill) and read once
        num += 8 * NS * 2 * plan.dense["gpool"]
    return sym, num
