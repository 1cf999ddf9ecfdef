"""Test-only torch restatement of the numeric phase of the star-mesh direct solver
(csrc/starmesh.cu) over the index lists of networksim_cntfet_b200/starmesh.py."""
import torch


def _segsum(values, ptr):
    """fixed-order segmented sums: values [NS, n], ptr [m+1] -> [NS, m]"""
    cs = torch.cat([torch.zeros((values.shape[0], 1), dtype=values.dtype, device=values.device), values.cumsum(1)], 1)
    return cs[:, ptr[1:]] - cs[:, ptr[:-1]]


def _segsum_exact(values, ptr):
    """segmented sums without cumulative cancellation (index_add on the segment id)"""
    m = ptr.numel() - 1
    seg = torch.repeat_interleave(torch.arange(m, device=values.device), ptr.diff())
    out = torch.zeros((values.shape[0], m), dtype=values.dtype, device=values.device)
    out.index_add_(1, seg, values)
    return out


def solve(plan, val, c, b):
    """val [NS, NNZ] (CSR off-diagonal values = -g), c [NS, R] coupling to the electrodes,
    b [NS, R] right-hand side; returns x [NS, R]"""
    NS = val.shape[0]
    g = _segsum_exact(-val[:, plan.e0_src], plan.e0_ptr)
    c = c.clone()
    b = b.clone()
    gE = torch.zeros((NS, max(plan.nnzL, 1)), dtype=val.dtype, device=val.device)
    dinvs = []
    for r in plan.rounds:
        ge = g[:, r.eslot]
        gE[:, r.gbase:r.gbase + ge.shape[1]] = ge
        d = c[:, r.E] + _segsum_exact(ge, r.nptr)
        dinv = 1.0 / d
        dinvs.append(dinv)
        gn = torch.zeros((NS, r.nE_new), dtype=val.dtype, device=val.device)
        has = r.copy_src >= 0
        gn[:, has] = g[:, r.copy_src[has]]
        if r.cm.numel():
            gn += _segsum_exact(g[:, r.ca] * g[:, r.cb] * dinv[:, r.cm], r.cptr)
        w = g[:, r.tslot] * dinv[:, r.tm]
        c[:, r.tn] += _segsum_exact(w * c[:, r.E][:, r.tm], r.tptr)
        b[:, r.tn] += _segsum_exact(w * b[:, r.E][:, r.tm], r.tptr)
        g = gn
    x = torch.zeros_like(b)
    if getattr(plan, "dense", None) is not None:
        _dense_tail(plan.dense, g, c, b, x)
    for r, dinv in zip(reversed(plan.rounds), reversed(dinvs)):
        ge = gE[:, r.gbase:r.gbase + r.eslot.numel()]
        x[:, r.E] = (b[:, r.E] + _segsum_exact(ge * x[:, r.nbr], r.nptr)) * dinv
    return x


def _dense_tail(d, g, c, b, x):
    """the unknowns left after the sparse rounds: the same star-mesh elimination on one dense
    upper-triangular conductance matrix per network, ascending rows (k_sm_dense)"""
    NS = g.shape[0]
    dptr, nodes, doff, epos = d["dptr"], d["dnodes"], d["doff"], d["epos"]
    pool = torch.zeros((NS, d["gpool"]), dtype=g.dtype, device=g.device)
    if d["nE"]:
        pool[:, epos] = g[:, :d["nE"]]
    for net in range(d["nnet"]):
        n = int(dptr[net + 1] - dptr[net])
        if n == 0:
            continue
        rows = nodes[int(dptr[net]):int(dptr[net + 1])]
        G = pool[:, int(doff[net]):int(doff[net]) + n * n].reshape(NS, n, n).clone()
        cc, bb = c[:, rows].clone(), b[:, rows].clone()
        dinv = torch.zeros((NS, n), dtype=g.dtype, device=g.device)
        for k in range(n):
            row = G[:, k, k + 1:]
            dinv[:, k] = 1.0 / (cc[:, k] + row.sum(1))
            f = row * dinv[:, k:k + 1]
            G[:, k + 1:, k + 1:] += torch.triu(f[:, :, None] * row[:, None, :], diagonal=1)
            cc[:, k + 1:] += f * cc[:, k:k + 1]
            bb[:, k + 1:] += f * bb[:, k:k + 1]
        xs = torch.zeros((NS, n), dtype=g.dtype, device=g.device)
        for k in range(n - 1, -1, -1):
            xs[:, k] = (bb[:, k] + (G[:, k, k + 1:] * xs[:, k + 1:]).sum(1)) * dinv[:, k]
        x[:, rows] = xs
