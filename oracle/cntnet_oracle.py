"""TEST INFRASTRUCTURE ONLY -- CPU (numpy / scipy) restatement of the reference hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module, and only as the checker or the timed CPU baseline.  The product
package `networksim_cntfet_b200` never imports it and has no CPU fallback.

Every function cites the reference lines (`/root/reference/<file>:<lines>`) it restates.
Semantics are the *canonical* ones of SURVEY.md section 0 (H1): nodes in ascending stick
index, 0.1 V source on stick 0, ground on stick 1.

Parity status: PINNED.  `tests/golden/*.npz` were produced by the unmodified reference
(`oracle/refshim.py`, `tests/golden/make_golden.py`) in the build container; the CPU test
suite checks this restatement against them (junction sets and cluster partition bit-exact,
V / |I| / I_source to 1e-12), and against the analytic 6-stick network of
`netsim.py:152-167`.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.csgraph as csgraph
import scipy.sparse.linalg as spla
import scipy.spatial as spatial

# stick kinds as u8 codes; 'v' = electrode, 'm' = metallic, 's' = semiconducting
KIND_V, KIND_M, KIND_S = 0, 1, 2
KIND_CHARS = "vms"
VDS = 0.1  # netsim.py:182, cnet.py:97

# gate rectangles [cx, cy, w, h] (netsim.py:274,276 ; measure_perc.py:180,184)
AREA_PARTIAL = (0.5, 0.0, 0.16, 0.667)
AREA_TOTAL = (0.217, 0.5, 0.167, 1.2)


def kinds_to_codes(kinds):
    """'v'/'m'/'s' strings -> u8 codes."""
    lut = {"v": KIND_V, "m": KIND_M, "s": KIND_S}
    return np.array([lut[k] for k in kinds], dtype=np.uint8)


def kind2_strings(kind2):
    """pair code 3*k1+k2 -> reference 2-char strings (netsim.py:148 `kinds[i]+kinds[j]`)."""
    return np.array([KIND_CHARS[c // 3] + KIND_CHARS[c % 3] for c in np.asarray(kind2)], dtype=object)


# --------------------------------------------------------------------------------------
# A2: stick generation  (netsim.py:97-125)
# --------------------------------------------------------------------------------------
def get_ends(xc, yc, angle, length):
    """netsim.py:97-103.  Returns xa, ya, xb, yb (endarray[0]=(x1,y1), endarray[1]=(x2,y2))."""
    xc, yc, angle, length = map(np.asarray, (xc, yc, angle, length))
    xa = xc - length / 2 * np.cos(angle)
    xb = xc + length / 2 * np.cos(angle)
    ya = yc + length / 2 * np.sin(angle)
    yb = yc - length / 2 * np.sin(angle)
    return xa, ya, xb, yb


def make_sticks_mt19937(n, l="exp", pm=0.135, scaling=5, seed=1):
    """netsim.py:105-125 with the legacy global MT19937 stream (`np.random.seed(seed)`,
    netsim.py:42).  Draw order per stick: pm-test, xc, yc, angle, [normal].  Returns the
    PRE-sort table (source first, drain last) as a dict of arrays."""
    rs = np.random.RandomState(seed)
    xc = np.empty(n + 2)
    yc = np.empty(n + 2)
    ang = np.empty(n + 2)
    ln = np.empty(n + 2)
    kind = np.empty(n + 2, dtype=np.uint8)
    xc[0], yc[0], ang[0], ln[0], kind[0] = 0.01, 0.5, np.pi / 2 - 1e-6, 100, KIND_V
    xc[-1], yc[-1], ang[-1], ln[-1], kind[-1] = 0.99, 0.5, np.pi / 2 - 1e-6, 100, KIND_V
    for i in range(1, n + 1):
        kind[i] = KIND_M if rs.rand() <= pm else KIND_S
        xc[i] = rs.rand()
        yc[i] = rs.rand()
        ang[i] = rs.rand() * 2 * np.pi
        if isinstance(l, str):
            ln[i] = abs(rs.normal(0.66, 0.44)) / scaling
        else:
            ln[i] = l / scaling
    return dict(xc=xc, yc=yc, angle=ang, length=ln, kind=kind)


def canonical_sort(t):
    """Canonical post-sort order used for GPU-generated sticks: source
    (x=0.01) at index 0, drain (x=0.99) at index 1, the rest by length descending, ties by
    original index (stable).  The reference's own order (pandas quicksort, netsim.py:133)
    differs only in which electrode is 0/1 and in tie order; parity tests feed the
    reference's post-sort table instead of calling this."""
    n = len(t["length"])
    body = np.arange(1, n - 1)
    order = body[np.argsort(-t["length"][body], kind="stable")]
    order = np.concatenate([[0, n - 1], order])
    out = {k: v[order] for k, v in t.items()}
    out["orig"] = order.astype(np.int64)  # reference: sticks['cluster']=index (netsim.py:132)
    return out


def sticks_with_ends(t):
    xa, ya, xb, yb = get_ends(t["xc"], t["yc"], t["angle"], t["length"])
    out = dict(t)
    out.update(xa=xa, ya=ya, xb=xb, yb=yb)
    return out


# --------------------------------------------------------------------------------------
# A3: intersection predicate (netsim.py:75-94), vectorised; same IEEE op sequence
# --------------------------------------------------------------------------------------
def check_intersect(x1a, y1a, x1b, y1b, x2a, y2a, x2b, y2b):
    """Returns (hit, xi, yi).  Stick 1 is the lower (longer) index, exactly as the
    reference calls it (netsim.py:146).  numpy float64 ufuncs never contract to FMA, so the
    rounding sequence equals the reference's scalar arithmetic."""
    with np.errstate(all="ignore"):
        # netsim.py:77-78 early-out: result-neutral, kept for fidelity
        early = (np.maximum(x1a, x1b) < np.minimum(x2a, x2b)) & (np.maximum(y1a, y1b) < np.minimum(y2a, y2b))
        m1 = (y1a - y1b) / (x1a - x1b)        # :81
        m2 = (y2a - y2b) / (x2a - x2b)        # :82
        b1 = y1a - m1 * x1a                   # :84
        b2 = y2a - m2 * x2a                   # :85
        par = m1 == m2                        # :86
        d = m1 - m2
        xi = (b2 - b1) / d                    # :89
        yi = (b2 * m1 - b1 * m2) / d          # :90
        inside = ((np.minimum(x1a, x1b) < xi) & (xi < np.maximum(x1a, x1b)) &
                  (np.minimum(x2a, x2b) < xi) & (xi < np.maximum(x2a, x2b)))  # :91
        hit = inside & ~par & ~early
    return hit, xi, yi


# --------------------------------------------------------------------------------------
# A4: junction finder (netsim.py:131-150)
# --------------------------------------------------------------------------------------
def _pairs_kdtree(xc, yc, length):
    """Candidate pairs i<j with |c_i-c_j| <= length_i (netsim.py:140-145)."""
    X = np.column_stack([xc, yc])
    tree = spatial.cKDTree(X)
    nb = tree.query_ball_point(X, length)
    cnt = np.fromiter((len(a) for a in nb), dtype=np.int64, count=len(nb))
    i = np.repeat(np.arange(len(nb)), cnt)
    j = np.concatenate([np.asarray(a, dtype=np.int64) for a in nb]) if len(nb) else np.zeros(0, np.int64)
    keep = i < j
    return i[keep], j[keep]


def _pairs_brute(xc, yc, length):
    """Same candidate set without the KD-tree: squared centre distance <= length_i^2."""
    n = len(xc)
    i, j = np.triu_indices(n, 1)
    dx = xc[i] - xc[j]
    dy = yc[i] - yc[j]
    keep = dx * dx + dy * dy <= length[i] * length[i]
    return i[keep], j[keep]


def find_junctions(s, pruner="kdtree"):
    """s: dict with xa,ya,xb,yb,xc,yc,length,kind in POST-sort order.
    Returns dict(stick1, stick2, x, y, kind2, ntests) sorted by (stick1, stick2); the
    reference's row order is KD-tree traversal order (an implementation detail, H4)."""
    if pruner == "kdtree":
        i, j = _pairs_kdtree(s["xc"], s["yc"], s["length"])
    else:
        i, j = _pairs_brute(s["xc"], s["yc"], s["length"])
    hit, xi, yi = check_intersect(s["xa"][i], s["ya"][i], s["xb"][i], s["yb"][i],
                                  s["xa"][j], s["ya"][j], s["xb"][j], s["yb"][j])
    with np.errstate(all="ignore"):
        hit &= (0 <= xi) & (xi <= 1) & (0 <= yi) & (yi <= 1)   # netsim.py:147
    i, j, xi, yi = i[hit], j[hit], xi[hit], yi[hit]
    order = np.lexsort((j, i))
    i, j, xi, yi = i[order], j[order], xi[order], yi[order]
    kind2 = (3 * s["kind"][i].astype(np.int64) + s["kind"][j]).astype(np.uint8)
    return dict(stick1=i.astype(np.int32), stick2=j.astype(np.int32), x=xi, y=yi, kind2=kind2,
                ntests=int(len(hit)))


# --------------------------------------------------------------------------------------
# A5/A6: components, percolation, cluster statistics (netsim.py:171-206, measure_perc.py:76-80)
# --------------------------------------------------------------------------------------
def label_clusters(n, stick1, stick2, orig=None):
    """Canonical min-stick-index labels (isolated sticks label themselves) and the
    statistics of SURVEY.md H5.
      nclust_true / maxclust_true : #distinct labels / largest component
      maxclust_ref : measure_perc.py:78 `len(max(nx.connected_components(g)))` == size of
                     the FIRST component == the one holding the smallest stick index that
                     has a junction (0 if there is no junction -> reference's `except`)
      nclust_ref   : measure_perc.py:76 `len(sticks.cluster.drop_duplicates())` where
                     sticks with a junction carry the component ordinal (components are
                     discovered in order of their minimum stick index) and isolated sticks
                     keep their PRE-sort index `orig` (netsim.py:132,197-206)."""
    stick1 = np.asarray(stick1, dtype=np.int64)
    stick2 = np.asarray(stick2, dtype=np.int64)
    g = sp.coo_matrix((np.ones(len(stick1)), (stick1, stick2)), shape=(n, n))
    _, comp = csgraph.connected_components(g, directed=False)
    minidx = np.full(comp.max() + 1, n, dtype=np.int64)
    np.minimum.at(minidx, comp, np.arange(n))
    label = minidx[comp].astype(np.int32)
    sizes = np.bincount(label, minlength=n)
    has_j = np.zeros(n, dtype=bool)
    has_j[stick1] = True
    has_j[stick2] = True
    out = dict(label=label, percolating=bool(n >= 2 and label[0] == label[1] and has_j[0]),
               nclust_true=int((sizes > 0).sum()), maxclust_true=int(sizes.max()) if n else 0)
    if has_j.any():
        first = int(np.flatnonzero(has_j)[0])
        out["maxclust_ref"] = int(sizes[label[first]])
    else:
        out["maxclust_ref"] = 0
    if orig is not None:
        roots = np.unique(label[has_j])                     # ascending min index == ordinal order
        ordinals = set(range(len(roots)))
        iso = set(np.asarray(orig)[~has_j].tolist())
        out["nclust_ref"] = len(ordinals | iso)
    return out


# --------------------------------------------------------------------------------------
# A7/A8: conductance elements (cnet.py:20-66)
# --------------------------------------------------------------------------------------
def _alpha_table(onoffmap):
    """cnet.py:23-27: alpha per pair code 3*k1+k2; nan where the reference has no key ('vv')."""
    a = np.full(9, np.nan)
    code = lambda s: 3 * KIND_CHARS.index(s[0]) + KIND_CHARS.index(s[1])
    base = {"ms": 0.5, "sm": 0.5, "mm": 1e-5, "ss": 1e-5, "vs": 1e-5, "sv": 1e-5, "vm": 1e-5, "mv": 1e-5}
    if onoffmap >= 1:
        base.update(vs=0.5, sv=0.5)
    if onoffmap >= 2:
        base.update(ss=0.5)
    if onoffmap not in (0, 1, 2):
        raise IndexError("onoffmap")
    for k, v in base.items():
        a[code(k)] = v
    return a


def conductance(kind2, vg, onoffmap=0, element="linexp"):
    """Per-junction conductance.  linexp: cnet.py:33-41 `exp(-alpha*vg)*exp(-10*alpha)`;
    fermidirac: cnet.py:52-66 `(1-offG)/(exp(10*(vg-0))+1)+offG`, offG = 5e-5 for ms/sm else 1."""
    kind2 = np.asarray(kind2)
    vg = np.broadcast_to(np.asarray(vg, dtype=np.float64), kind2.shape)
    if element == "linexp":
        alpha = _alpha_table(onoffmap)[kind2]
        normalization = np.exp(-10 * alpha)
        return np.exp(-alpha * vg) * normalization
    elif element == "fermidirac":
        if onoffmap != 0:
            raise IndexError("FermiDiracTransistor only defines onoffmap 0 (cnet.py:53)")
        ratio = np.ones(9)
        ratio[3 * KIND_M + KIND_S] = 1 / 5e-5
        ratio[3 * KIND_S + KIND_M] = 1 / 5e-5
        ratio[0] = np.nan
        offG = 1 / ratio[kind2]
        scaling = 1 - offG
        with np.errstate(over="ignore"):
            return scaling * (1 / (np.exp(10 * (vg - 0)) + 1)) + offG
    raise ValueError(element)


def in_area(x, y, area):
    """cnet.py:166-175 closed rectangle [cx +- w/2] x [cy +- h/2]."""
    right = area[0] + area[2] / 2
    left = area[0] - area[2] / 2
    top = area[1] + area[3] / 2
    bottom = area[1] - area[3] / 2
    return (left <= x) & (x <= right) & (bottom <= y) & (y <= top)


def gate_voltages(jx, jy, gate, vg):
    """Per-junction gate voltage after `RandomCNTNetwork.gate(vg, gate)` (netsim.py:266-277):
    everything reset to 0, then 'back' = all junctions, 'partial'/'total' = junctions inside
    the rectangle."""
    v = np.zeros(len(jx))
    if gate == "back":
        v[:] = vg
    elif gate == "partial":
        v[in_area(jx, jy, AREA_PARTIAL)] = vg
    elif gate == "total":
        v[in_area(jx, jy, AREA_TOTAL)] = vg
    elif gate is not None:
        raise ValueError(gate)
    return v


# --------------------------------------------------------------------------------------
# A9-A13: MNA assemble + solve (cnet.py:104-156) in the reduced SPD form of SURVEY.md 8(a)
# --------------------------------------------------------------------------------------
def assemble_reduced(n, stick1, stick2, g, label):
    """Reduced system L_UU V_U = 0.1 * g_{U,source} on the percolating component.
    Nodes ascending; source = stick 0, drain/ground = stick 1 (canonical H1).
    Returns A (csr), b, nodes (stick index of every unknown)."""
    stick1 = np.asarray(stick1, dtype=np.int64)
    stick2 = np.asarray(stick2, dtype=np.int64)
    comp = label == label[0]
    nodes_c = np.flatnonzero(comp)           # ascending; [0, 1, ...]
    assert nodes_c[0] == 0 and nodes_c[1] == 1
    cid = np.full(n, -1, dtype=np.int64)
    cid[nodes_c] = np.arange(len(nodes_c))
    e = comp[stick1]                         # both ends share the component
    a, b_, ge = cid[stick1[e]], cid[stick2[e]], np.asarray(g)[e]
    nc = len(nodes_c)
    W = sp.coo_matrix((np.concatenate([ge, ge]), (np.concatenate([a, b_]), np.concatenate([b_, a]))),
                      shape=(nc, nc)).tocsr()
    deg = np.asarray(W.sum(axis=1)).ravel()
    L = (sp.diags(deg) - W).tocsr()
    A = L[2:, 2:].tocsr()
    rhs = VDS * np.asarray(W[2:, 0].todense()).ravel()
    return A, rhs, nodes_c[2:], dict(W=W, nodes_c=nodes_c, edge_mask=e)


def _refine(A, rhs, x, steps):
    """Iterative refinement with extended-precision residuals: SuperLU alone leaves up to
    ~5e-9 relative error at Vg=+10 on 100x100 um networks (so does the reference's own solve);
    a few refinement steps bring the checker to ~1e-15."""
    A = A.tocsr()
    lu = spla.splu(A.tocsc())
    rows = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
    dv = A.data.astype(np.longdouble)
    xl = x.astype(np.longdouble)
    for _ in range(steps):
        Ax = np.zeros(A.shape[0], dtype=np.longdouble)
        np.add.at(Ax, rows, dv * xl[A.indices])
        res = rhs.astype(np.longdouble) - Ax
        xl = xl + lu.solve(res.astype(np.float64)).astype(np.longdouble)
    return xl.astype(np.float64)


def solve_network(n, stick1, stick2, g, label, refine=0):
    """Direct solve (the reference uses SuperLU `spsolve`, cnet.py:147-149), optionally
    followed by `refine` steps of iterative refinement.
    Returns dict(V[n] (nan outside the component), I_edge[J] (nan outside), I_source)."""
    A, rhs, nodes_u, aux = assemble_reduced(n, stick1, stick2, g, label)
    vu = spla.spsolve(A.tocsc(), rhs) if A.shape[0] else np.zeros(0)
    if refine and A.shape[0]:
        vu = _refine(A, rhs, vu, refine)
    V = np.full(n, np.nan)
    V[0] = VDS
    V[1] = 0.0
    V[nodes_u] = vu
    stick1 = np.asarray(stick1, dtype=np.int64)
    stick2 = np.asarray(stick2, dtype=np.int64)
    g = np.asarray(g)
    I = np.abs(g * (V[stick1] - V[stick2]))         # cnet.py:140-146
    e = aux["edge_mask"]
    src = (stick1 == 0) & e                         # stick 0 is always stick1
    drn = (stick1 == 1) & e
    # The MNA unknown x[-1] (cnet.py:135) is the current through the source.  Three
    # algebraically equal estimators; in floating point the source-side sum cancels badly
    # (neighbours of the source sit within ~1e-6 of 0.1 V, so FP64 *representation* of V
    # already costs 1e-8 relative at 100x100 um) while dissipated power / Vds is a sum of
    # positive terms whose error is second order in the voltage error.  I_source is the
    # power form; the other two are returned for diagnostics.
    dV = V[stick1[e]] - V[stick2[e]]
    I_power = float(np.sum(g[e] * dV * dV) / VDS)
    I_src_side = float(np.sum(g[src] * (VDS - V[stick2[src]])))
    I_drn_side = float(np.sum(g[drn] * V[stick2[drn]]))
    I_source = I_power
    return dict(V=V, I=I, I_source=I_source, I_src_side=I_src_side, I_drn_side=I_drn_side,
                A=A, rhs=rhs, nodes_u=nodes_u)


def solve_network_exact(n, stick1, stick2, g, label, dtype=np.longdouble):
    """Voltages of the IDEAL network (exact-arithmetic conductances, no assembled matrix): sequential
    minimum-degree star-mesh elimination in extended precision.  Every quantity is a sum of positive
    terms, so the result is accurate to a few ulp of `dtype` whatever the conditioning -- unlike
    `solve_network`, whose answer (refined or not) is the solution of the ASSEMBLED FP64 matrix and
    inherits the rounding of its diagonal (1e-8 Vds on back-gated 100x100 um networks).  Same physics
    as cnet.py:104-149 (source = stick 0 at 0.1 V, ground = stick 1).  Returns V[n] (nan outside the
    percolating component).  Pure Python loops: meant for checks, not for timing."""
    import heapq
    stick1 = np.asarray(stick1, dtype=np.int64)
    stick2 = np.asarray(stick2, dtype=np.int64)
    g = np.asarray(g)
    comp = label == label[0]
    V = np.full(n, np.nan)
    V[0], V[1] = VDS, 0.0
    adj = {int(i): {} for i in np.flatnonzero(comp) if i >= 2}
    c = {i: dtype(0) for i in adj}          # conductance to the fixed potentials
    b = {i: dtype(0) for i in adj}          # sum of conductance x fixed potential
    for a, w, ge in zip(stick1.tolist(), stick2.tolist(), g.tolist()):
        if not comp[a]:
            continue
        ge = dtype(ge)
        if a < 2:                           # stick 0 / 1 is always stick1 of its junctions
            c[w] += ge
            b[w] += ge * dtype(VDS if a == 0 else 0.0)
        else:
            adj[a][w] = adj[a].get(w, dtype(0)) + ge
            adj[w][a] = adj[w].get(a, dtype(0)) + ge
    heap = [(len(nb), i) for i, nb in adj.items()]
    heapq.heapify(heap)
    order, done = [], set()
    while heap:
        d, v = heapq.heappop(heap)
        if v in done or d != len(adj[v]):
            continue                        # stale entry
        done.add(v)
        nb = adj[v]
        dv = c[v] + sum(nb.values(), dtype(0))
        order.append((v, dict(nb), dv))
        items = list(nb.items())
        for u, gu in items:
            del adj[u][v]
            f = gu / dv
            c[u] += f * c[v]
            b[u] += f * b[v]
            for w, gw in items:
                if w != u:
                    adj[u][w] = adj[u].get(w, dtype(0)) + f * gw
        for u, _ in items:
            heapq.heappush(heap, (len(adj[u]), u))
    x = {}
    for v, nb, dv in reversed(order):
        x[v] = (b[v] + sum((gu * x[u] for u, gu in nb.items()), dtype(0))) / dv
    for v, xv in x.items():
        V[v] = float(xv)
    return V


# --------------------------------------------------------------------------------------
# A17: measure_fullnet (measure_perc.py:140-203), numeric columns only
# --------------------------------------------------------------------------------------
def measure_fullnet(s, onoffmap=1, element="linexp", junctions=None, refine=0):
    """s = post-sort stick dict.  Returns dict(nclust*, maxclust*, percolating, ion,
    ioff_back, ioff_total, ioff_partial)."""
    n = len(s["xc"])
    j = junctions if junctions is not None else find_junctions(s)
    c = label_clusters(n, j["stick1"], j["stick2"], s.get("orig"))
    out = {k: v for k, v in c.items() if k != "label"}
    out.update(njunctions=len(j["stick1"]), ntests=j["ntests"], ion=0.0, ioff_back=0.0,
               ioff_total=0.0, ioff_partial=0.0)
    if c["percolating"]:
        for name, gate, vg in (("ion", None, 0.0), ("ioff_back", "back", 10.0),
                               ("ioff_total", "total", 10.0), ("ioff_partial", "partial", 10.0)):
            vgj = gate_voltages(j["x"], j["y"], gate, vg)
            g = conductance(j["kind2"], vgj, onoffmap, element)
            out[name] = solve_network(n, j["stick1"], j["stick2"], g, c["label"], refine=refine)["I_source"]
    return out
