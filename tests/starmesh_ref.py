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
    for r, dinv in zip(reversed(plan.rounds), reversed(dinvs)):
        ge = gE[:, r.gbase:r.gbase + r.eslot.numel()]
        x[:, r.E] = (b[:, r.E] + _segsum_exact(ge * x[:, r.nbr], r.nptr)) * dinv
    return x
