"""Test-only numpy restatement of the star-mesh direct solver: the symbolic phase of
csrc/starmesh_sym.cu (same rounds, same slot numbering, same list layout -- the CUDA plan must equal
this one array by array) and the numeric phase of csrc/starmesh.cu over those lists."""
import numpy as np

SAT = 1 << 29


class RefPlan(object):
    pass


def _priority(deg, lid):
    h = ((np.uint64(lid) * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)) >> np.uint64(12)
    d = np.uint64(min(int(deg), (1 << 18) - 1))
    return int((d << np.uint64(46)) | (h << np.uint64(26)) | np.uint64(lid))


def symbolic(roff, rowptr, col, dense_max=0):
    """roff [B+1], rowptr [R+1], col [NNZ] (network-local, ascending inside a row; symmetric, no diagonal).
    Returns a RefPlan whose arrays are what cnt_sm_symbolic writes."""
    roff = np.asarray(roff, dtype=np.int64)
    rowptr = np.asarray(rowptr, dtype=np.int64)
    col = np.asarray(col, dtype=np.int64)
    B, R = len(roff) - 1, int(roff[-1])
    rnet = np.searchsorted(roff[1:], np.arange(R), side="right")
    # round-0 slots: unique upper pairs, row-major
    eu, ew, e0_ptr, e0_src = [], [], [], []
    deg = np.zeros(R, dtype=np.int64)
    slot_of = {}
    for i in range(R):
        base = roff[rnet[i]]
        prev = prevu = -1
        for k in range(rowptr[i], rowptr[i + 1]):
            c = int(col[k] + base)
            if c == i:
                continue
            if c != prev:
                deg[i] += 1
            prev = c
            if c < i:
                continue
            if c != prevu:
                slot_of[(i, c)] = len(eu)
                eu.append(i)
                ew.append(c)
                e0_ptr.append(len(e0_src))
                prevu = c
            e0_src.append(k)
    e0_ptr.append(len(e0_src))
    nE0 = len(eu)
    mid = np.full(R, -1, dtype=np.int64)
    nalive = np.diff(roff).copy()
    live = list(range(nE0))
    nodes = list(range(R))
    piv, nptr, nbr, eslot = [], [0], [], []
    tgt, cptr, cm, ca, cb = [], [], [], [], []
    tn, tptr, tm, tslot = [], [], [], []
    rounds = []
    max_deg = 0
    while True:
        frozen = nalive <= dense_max
        blocked = set()
        for e in live:
            u, w = eu[e], ew[e]
            net = rnet[u]
            if frozen[net]:
                continue
            pu = _priority(deg[u], u - roff[net])
            pw = _priority(deg[w], w - roff[net])
            blocked.add(u if pu > pw else w)
        P = [v for v in nodes if not frozen[rnet[v]] and v not in blocked]
        if not P:
            break
        survivors = [v for v in nodes if not frozen[rnet[v]] and v in blocked]
        ebase, lbase, cbase, gebase, gnbase, sbase = len(piv), len(nbr), len(cm), len(tgt), len(tn), len(eu)
        for ml, v in enumerate(P):
            mid[v] = ebase + ml
        inc = {v: [] for v in P}
        newlive = []
        for e in live:
            u, w = eu[e], ew[e]
            if frozen[rnet[u]]:
                continue
            if mid[u] >= ebase:
                inc[u].append((w, e))
            elif mid[w] >= ebase:
                inc[w].append((u, e))
            else:
                newlive.append(e)
        pairs = []       # (m, list position of edge i, of edge j, key)
        for ml, v in enumerate(P):
            lst = sorted(inc[v])
            k0 = len(nbr)
            piv.append(v)
            max_deg = max(max_deg, len(lst))
            for o, e in lst:
                nbr.append(o)
                eslot.append(e)
                deg[o] -= 1
            nptr.append(len(nbr))
            nalive[rnet[v]] -= 1
            for i in range(len(lst)):
                for j in range(i + 1, len(lst)):
                    pairs.append((ebase + ml, k0 + i, k0 + j, (lst[i][0], lst[j][0])))
        # edge-target groups in order of their first pair
        groups = {}
        for t, (m, sa, sb, key) in enumerate(pairs):
            groups.setdefault(key, []).append(t)
        nnew = 0
        for key, members in groups.items():          # dicts keep insertion order = order of the first pair
            if key in slot_of:
                s = slot_of[key]
            else:
                s = len(eu)
                slot_of[key] = s
                eu.append(key[0])
                ew.append(key[1])
                deg[key[0]] += 1
                deg[key[1]] += 1
                newlive.append(s)
                nnew += 1
            tgt.append(s)
            cptr.append(len(cm))
            for t in members:
                cm.append(pairs[t][0])
                ca.append(pairs[t][1])
                cb.append(pairs[t][2])
        # node-target groups in order of their first incident entry
        ngroups = {}
        for k in range(lbase, len(nbr)):
            ngroups.setdefault(nbr[k], []).append(k)
        lm = np.repeat(np.arange(ebase, len(piv)), np.diff(nptr[ebase:]))
        for o, members in ngroups.items():
            tn.append(o)
            tptr.append(len(tm))
            for k in members:
                tm.append(int(lm[k - lbase]))
                tslot.append(k)
        rounds.append(dict(nelim=len(P), nlist=len(nbr) - lbase, npairs=len(pairs), nge=len(groups), ngn=len(ngroups),
                           nnew=nnew, ebase=ebase, lbase=lbase, cbase=cbase, gebase=gebase, gnbase=gnbase, sbase=sbase,
                           nlive=len(live), nnodes=len(nodes)))
        live = newlive
        nodes = survivors
    cptr.append(len(cm))
    tptr.append(len(tm))
    p = RefPlan()
    i32 = lambda a: np.asarray(a, dtype=np.int64)
    p.R, p.B, p.nE0, p.nslots, p.rounds, p.max_deg = R, B, nE0, len(eu), rounds, max_deg
    p.piv, p.nptr, p.nbr, p.eslot = i32(piv), i32(nptr), i32(nbr), i32(eslot)
    p.tgt, p.cptr, p.cm, p.ca, p.cb = i32(tgt), i32(cptr), i32(cm), i32(ca), i32(cb)
    p.tn, p.tptr, p.tm, p.tslot = i32(tn), i32(tptr), i32(tm), i32(tslot)
    p.e0_ptr, p.e0_src = i32(e0_ptr), i32(e0_src)
    p.eu, p.ew = i32(eu), i32(ew)
    # dense tail
    alive = mid < 0
    p.dnodes = np.flatnonzero(alive)
    cnt = nalive.astype(np.int64)
    assert cnt.sum() == len(p.dnodes)
    p.dptr = np.r_[0, np.cumsum(cnt)]
    p.doff = np.r_[0, np.cumsum(cnt * cnt)]
    loc = np.full(R, -1, dtype=np.int64)
    loc[p.dnodes] = np.arange(len(p.dnodes)) - p.dptr[rnet[p.dnodes]]
    ds, dp = [], []
    for e in range(len(eu)):
        u, w = eu[e], ew[e]
        if alive[u] and alive[w]:
            net = rnet[u]
            ds.append(e)
            dp.append(p.doff[net] + loc[u] * cnt[net] + loc[w])
    p.dslot, p.depos = i32(ds), i32(dp)
    p.dedge = np.r_[0, np.cumsum(np.bincount(rnet[p.eu[p.dslot]], minlength=B))] if len(ds) else np.zeros(B + 1, dtype=np.int64)
    p.dense_nmax = int(cnt.max()) if B else 0
    p.dense_gpool = int(p.doff[-1])
    p.nnzL = len(nbr)
    return p


def _segsum(values, ptr):
    """sequential sums of values [NS, n] over the segments ptr [m+1] -> [NS, m]"""
    m = len(ptr) - 1
    seg = np.repeat(np.arange(m), np.diff(ptr))
    return np.stack([np.bincount(seg, weights=v, minlength=m) for v in values])


def solve(plan, val, c, b):
    """val [NS, NNZ] (CSR off-diagonal values = -g), c [NS, R] coupling to the electrodes,
    b [NS, R] right-hand side; returns x [NS, R]"""
    val, c, b = np.asarray(val, dtype=np.float64), np.array(c, dtype=np.float64), np.array(b, dtype=np.float64)
    NS = val.shape[0]
    g = np.zeros((NS, max(plan.nslots, 1)))
    g[:, :plan.nE0] = _segsum(-val[:, plan.e0_src], plan.e0_ptr)
    dinv = np.zeros((NS, plan.R))
    gl = np.zeros((NS, max(len(plan.nbr), 1)))
    for r in plan.rounds:
        ms = np.arange(r["ebase"], r["ebase"] + r["nelim"])
        ptr = plan.nptr[r["ebase"]:r["ebase"] + r["nelim"] + 1]
        ge = g[:, plan.eslot[ptr[0]:ptr[-1]]]
        gl[:, ptr[0]:ptr[-1]] = ge                  # dead edges in list order: what ca / cb / tslot index
        d = c[:, plan.piv[ms]] + _segsum(ge, ptr - ptr[0])
        dinv[:, ms] = 1.0 / d
        if r["nge"]:
            gs = slice(r["gebase"], r["gebase"] + r["nge"])
            cp = plan.cptr[r["gebase"]:r["gebase"] + r["nge"] + 1]
            ks = slice(cp[0], cp[-1])
            add = _segsum(gl[:, plan.ca[ks]] * gl[:, plan.cb[ks]] * dinv[:, plan.cm[ks]], cp - cp[0])
            g[:, plan.tgt[gs]] += add
        if r["ngn"]:
            gs = slice(r["gnbase"], r["gnbase"] + r["ngn"])
            tp = plan.tptr[r["gnbase"]:r["gnbase"] + r["ngn"] + 1]
            ks = slice(tp[0], tp[-1])
            w = gl[:, plan.tslot[ks]] * dinv[:, plan.tm[ks]]
            pv = plan.piv[plan.tm[ks]]
            c[:, plan.tn[gs]] += _segsum(w * c[:, pv], tp - tp[0])
            b[:, plan.tn[gs]] += _segsum(w * b[:, pv], tp - tp[0])
    x = np.zeros_like(b)
    _dense_tail(plan, g, c, b, x)
    for r in reversed(plan.rounds):
        ms = np.arange(r["ebase"], r["ebase"] + r["nelim"])
        ptr = plan.nptr[r["ebase"]:r["ebase"] + r["nelim"] + 1]
        ks = slice(ptr[0], ptr[-1])
        v = plan.piv[ms]
        x[:, v] = (b[:, v] + _segsum(gl[:, ks] * x[:, plan.nbr[ks]], ptr - ptr[0])) * dinv[:, ms]
    return x


def _dense_tail(plan, g, c, b, x):
    """the unknowns left after the sparse rounds: the same star-mesh elimination on one dense
    upper-triangular conductance matrix per network, ascending rows (k_sm_dense)"""
    NS = g.shape[0]
    pool = np.zeros((NS, max(plan.dense_gpool, 1)))
    if len(plan.dslot):
        pool[:, plan.depos] = g[:, plan.dslot]
    for net in range(plan.B):
        n = int(plan.dptr[net + 1] - plan.dptr[net])
        if n == 0:
            continue
        rows = plan.dnodes[plan.dptr[net]:plan.dptr[net + 1]]
        G = pool[:, plan.doff[net]:plan.doff[net] + n * n].reshape(NS, n, n).copy()
        cc, bb = c[:, rows].copy(), b[:, rows].copy()
        dinv = np.zeros((NS, n))
        for k in range(n):
            row = G[:, k, k + 1:]
            dinv[:, k] = 1.0 / (cc[:, k] + row.sum(1))
            f = row * dinv[:, k:k + 1]
            G[:, k + 1:, k + 1:] += np.triu(f[:, :, None] * row[:, None, :], k=1)
            cc[:, k + 1:] += f * cc[:, k:k + 1]
            bb[:, k + 1:] += f * bb[:, k:k + 1]
        xs = np.zeros((NS, n))
        for k in range(n - 1, -1, -1):
            xs[:, k] = (bb[:, k] + (G[:, k, k + 1:] * xs[:, k + 1:]).sum(1)) * dinv[:, k]
        x[:, rows] = xs
