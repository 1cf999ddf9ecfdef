"""Test-only numpy restatements of the two exact reductions the GPU applies before assembly
(csrc/prune.cu): leaf pruning to the 2-core and series elimination of an independent set of
degree-2 sticks.  Used by the GPU parity tests (as the checker of cnt_prune_leaves /
cnt_series_select) and by the CPU tests that validate the reductions themselves against the
reference goldens."""
import numpy as np


def two_core(n, s1, s2, comp):
    """numpy peeling: sticks of the component left after removing dangling sticks repeatedly
    (electrodes 0 and 1 are never removed); returns (alive mask, attachment stick or -1)"""
    alive = comp.copy()
    attach = np.full(n, -1, np.int64)
    parent = np.full(n, -1, np.int64)
    while True:
        m = alive[s1] & alive[s2]
        deg = np.bincount(np.r_[s1[m], s2[m]], minlength=n)
        leaf = alive & (deg == 1) & (np.arange(n) >= 2)
        if not leaf.any():
            break
        for a, c in ((s1[m], s2[m]), (s2[m], s1[m])):
            k = leaf[a]
            parent[a[k]] = c[k]
        # two leaves joined to each other cannot occur inside a percolating component
        alive &= ~leaf
    for s in np.nonzero(parent >= 0)[0]:
        p = parent[s]
        while not alive[p]:
            p = parent[p]
        attach[s] = p
    return alive, attach


def series_set(n, s1, s2, alive):
    """numpy restatement of cnt_series_select: candidates are the non-electrode sticks of the
    2-core with exactly two junctions inside it, both to non-electrode sticks; a candidate is
    eliminated unless one of its two neighbours is a candidate with a smaller index."""
    m = alive[s1] & alive[s2]
    a, c, e = s1[m], s2[m], np.nonzero(m)[0]
    deg = np.bincount(np.r_[a, c], minlength=n)
    nb = [[] for _ in range(n)]
    for x, y, k in zip(np.r_[a, c].tolist(), np.r_[c, a].tolist(), np.r_[e, e].tolist()):
        if deg[x] == 2:
            nb[x].append((y, k))
    cand = np.zeros(n, bool)
    for s in range(2, n):
        if alive[s] and deg[s] == 2 and nb[s][0][0] >= 2 and nb[s][1][0] >= 2:
            cand[s] = True
    out = {}
    for s in np.nonzero(cand)[0].tolist():
        (n0, e0), (n1, e1) = sorted(nb[s])
        if not (cand[n0] and n0 < s) and not (cand[n1] and n1 < s):
            out[s] = (n0, n1, e0, e1)
    return out


