"""Test-only torch restatement of the per-rank phases of csrc/pcg_slab.cu (same SlabState tensors,
same schedule) -- lets the partition plan and the exchange schedule of
networksim_cntfet_b200/slab.py be checked on CPU (LocalComm in-process, DistComm over gloo) and
serves as the checker of the CUDA phases on the GPU.  Never imported by the package."""
import numpy as np
import torch

from networksim_cntfet_b200 import _lib

RZ, BB, RR, BETA, ITERS, STATUS = range(6)


class TorchPhases(object):
    @staticmethod
    def block_rows():
        return 256

    def _sums(self, st):
        n = st.n_own
        for q in range(st.NS):
            r = st.r[q, :n]
            nl, ns = int(st.nloc[q]), int(st.nsh[q])
            ra = st.row_agg[q, :n].long()
            S = torch.zeros(max(nl, 1), dtype=torch.float64, device=r.device)
            m = ra >= 0
            S.index_add_(0, ra[m], r[m])
            st.S[q, :nl] = S[:nl]
            local_only = st.agg_shared[q, :nl] < 0
            e = (st.einv[q, :nl] * S[:nl] ** 2)[local_only].sum()
            st.red2[q, 0] = (r * r * st.dinv[q, :n]).sum() + e
            st.red2[q, 1] = (r * r).sum()
            st.red2[q, 2:] = 0.0
            sl = st.shared_local[q, :ns].long()
            has = sl >= 0
            vals = torch.zeros(ns, dtype=torch.float64, device=r.device)
            vals[has] = S[sl[has]]
            st.red2[q, 2:2 + ns] = vals

    def init(self, st):
        n = st.n_own
        st.done.zero_()
        st.p.zero_()
        st.x.zero_()
        st.dinv[:, :n] = 1.0 / st.diag[:, :n]
        st.r[:, :n] = st.rhs[:, :n]
        self._sums(st)

    def update(self, st):
        n = st.n_own
        for q in range(st.NS):
            if int(st.done[q]):
                continue
            pq = float(st.red_pq[q])
            alpha = float(st.scal[q, RZ]) / pq if pq > 0 else 0.0
            st.x[q, :n] += alpha * st.p[q, :n]
            st.r[q, :n] -= alpha * st.q[q, :n]
        self._sums(st)

    def direction(self, st, first):
        n = st.n_own
        for q in range(st.NS):
            nl, ns = int(st.nloc[q]), int(st.nsh[q])
            v = st.red2[q, 2:2 + ns]
            rz = float(st.red2[q, 0] + (st.einv_sh[q, :ns] * v * v).sum())
            rr = float(st.red2[q, 1])
            sl = st.shared_local[q, :ns].long()
            has = sl >= 0
            st.S[q, sl[has]] = v[has]
            sc = st.scal[q]
            if first:
                sc[RZ], sc[BB], sc[RR], sc[BETA], sc[ITERS], sc[STATUS] = rz, rr, rr, 0.0, 0.0, 0.0
                st.done[q] = 1 if rr == 0.0 else 0
            elif int(st.done[q]):
                sc[BETA] = 0.0
            elif not float(st.red_pq[q]) > 0.0:
                st.done[q] = 1
                sc[STATUS] = 2.0
                sc[BETA] = 0.0
            else:
                sc[BETA] = rz / float(sc[RZ])
                sc[RZ], sc[RR] = rz, rr
                sc[ITERS] += 1.0
                if rr <= st.rtol * st.rtol * float(sc[BB]):
                    st.done[q] = 1
                elif float(sc[ITERS]) >= st.maxit:
                    st.done[q] = 1
                    sc[STATUS] = 1.0
            ra = st.row_agg[q, :n].long()
            z = st.r[q, :n] * st.dinv[q, :n]
            m = ra >= 0
            z[m] += st.einv[q, ra[m]] * st.S[q, ra[m]]
            beta = float(sc[BETA])
            st.p[q, :n] = z + beta * st.p[q, :n] if beta != 0.0 else z
            st.sendbuf[q, :st.n_exp] = st.p[q, st.exp_rows.long()]

    def spmv(self, st):
        n = st.n_own
        flat = st.recvbuf.reshape(-1)
        for q in range(st.NS):
            if st.n_halo:
                st.p[q, n:n + st.n_halo] = flat[st.halo_src + q * st.emax]
            if int(st.done[q]):
                st.red_pq[q] = 0.0
                continue
            rows = torch.repeat_interleave(torch.arange(n, device=st.p.device), st.rowptr.long().diff())
            acc = st.diag[q, :n] * st.p[q, :n]
            acc = acc.index_add(0, rows, st.val[q] * st.p[q, st.col.long()])
            st.q[q, :n] = acc
            st.red_pq[q] = (st.p[q, :n] * acc).sum()


def strong_aggregates(A, theta=0.1):
    """numpy/scipy: aggregates = connected components (>= 2 rows) of the entries with
    |a_ij| > theta * max|a_ij|; returns (agg_of_row [-1 = none], einv)"""
    import scipy.sparse as sp
    from scipy.sparse.csgraph import connected_components
    A = A.tocsr()
    off = (A - sp.diags(A.diagonal())).tocoo()
    strong = np.abs(off.data) > theta * np.abs(off.data).max()
    G = sp.coo_matrix((np.ones(strong.sum()), (off.row[strong], off.col[strong])), shape=A.shape)
    nc, lab = connected_components(G, directed=False)
    size = np.bincount(lab, minlength=nc)
    keep = size[lab] >= 2
    ids = np.full(nc, -1, np.int64)
    ids[np.unique(lab[keep])] = np.arange(len(np.unique(lab[keep])))
    agg = np.where(keep, ids[lab], -1)
    na = int(agg.max()) + 1 if keep.any() else 0
    Z = sp.coo_matrix((np.ones(keep.sum()), (np.nonzero(keep)[0], agg[keep])), shape=(A.shape[0], max(na, 1))).tocsr()
    E = (Z.T @ A @ Z).diagonal()
    einv = np.where(E > 0, 1.0 / np.where(E > 0, E, 1.0), 0.0)
    return agg.astype(np.int32), einv[:max(na, 1)]


def random_network_system(n=40, seed=0, nconf=2):
    """SPD test systems on a jittered grid graph with bimodal conductances (strong 1, weak e^-5 ..
    e^-10) and Dirichlet couplings on two sides: (rowptr, col, val[NS], diag[NS], rhs[NS]) as numpy,
    rows in Z-order-like (row-major) numbering, plus the scipy matrices."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    N = n * n
    idx = np.arange(N).reshape(n, n)
    e = np.r_[np.c_[idx[:, :-1].ravel(), idx[:, 1:].ravel()], np.c_[idx[:-1, :].ravel(), idx[1:, :].ravel()],
              np.c_[idx[:-1, :-1].ravel(), idx[1:, 1:].ravel()][rng.random((n - 1) * (n - 1)) < 0.3]]
    mats, vals, diags, rhss = [], [], [], []
    pattern = None
    for c in range(nconf):
        weak = rng.random(len(e)) < 0.35
        g = np.where(weak, np.exp(-5.0 - 5.0 * (c % 2)), 1.0) * (1 + 0.01 * rng.random(len(e)))
        Wm = sp.coo_matrix((np.r_[g, g], (np.r_[e[:, 0], e[:, 1]], np.r_[e[:, 1], e[:, 0]])), shape=(N, N)).tocsr()
        gs = np.zeros(N)
        gd = np.zeros(N)
        gs[idx[:, 0]] = np.exp(-5.0)
        gd[idx[:, -1]] = np.exp(-5.0)
        d = np.asarray(Wm.sum(1)).ravel() + gs + gd
        A = (sp.diags(d) - Wm).tocsr()
        A.sort_indices()
        offd = (-Wm).tocsr()
        offd.sort_indices()
        if pattern is None:
            pattern = (offd.indptr.astype(np.int32), offd.indices.astype(np.int32))
        assert np.array_equal(pattern[1], offd.indices)
        mats.append(A)
        vals.append(offd.data.copy())
        diags.append(d)
        rhss.append(0.1 * gs)
    return pattern[0], pattern[1], np.stack(vals), np.stack(diags), np.stack(rhss), mats


def to_states(system, comm, phases, device="cpu", theta=0.1, rtol=1e-13, maxit=20000, precond=True):
    from networksim_cntfet_b200 import slab
    rowptr, col, val, diag, rhs, mats = system
    NS, R = diag.shape
    t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=device)
    agg_of_row = agg_einv = None
    if precond:
        aggs, einvs, base = [], [], 0
        for A in mats:
            a, e = strong_aggregates(A, theta)
            aggs.append(np.where(a >= 0, a + base, -1))
            einvs.append(e)
            base += len(e)
        agg_of_row, agg_einv = t(np.concatenate(aggs), torch.int32), t(np.concatenate(einvs), torch.float64)
    return slab.build_states(t(rowptr, torch.int32), t(col, torch.int32), t(val, torch.float64), t(diag, torch.float64),
                             t(rhs, torch.float64), agg_of_row, agg_einv, comm, phases.block_rows(), rtol, maxit)
