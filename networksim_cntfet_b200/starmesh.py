"""Direct solve of the junction-graph systems by parallel star-mesh elimination (host driver).

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
100x100 um system (46k unknowns) creates only ~1.6x fill and ~1e6 multiply-adds, against 1.7e8 row
updates of the 3.7k PCG iterations.

Both phases are hand-written CUDA behind the C-ABI (include/cntnet_b200.h section 5c):

* `cnt_sm_symbolic` (csrc/starmesh_sym.cu) -- topology only, once per batch pattern: every round
  eliminates the independent set of unknowns with the locally smallest (degree, hash, index) for
  all networks at once; edge slots never move, an insert-only hash finds the target of every
  star-mesh contribution; one read-back of the plan header per batch.
* `cnt_sm_solve` (csrc/starmesh.cu) -- all gate configurations at once, fixed-order sums.

This module only sizes the buffers (PyTorch tensors as device memory) and calls the two.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check

MAX_ROUNDS = 1024
ARRAYS = ("piv", "nptr", "nbr", "eslot", "tgt", "cptr", "cm", "ca", "cb", "tn", "tptr", "tm", "tslot", "e0_ptr", "e0_src",
          "dptr", "dnodes", "dslot", "doff", "depos", "info", "dedge")


class SmRound(C.Structure):
    """mirror of cnt_sm_round (include/cntnet_b200.h)"""
    _fields_ = [(n, C.c_int32) for n in ("nelim", "nlist", "npairs", "nge", "ngn", "nnew", "ebase", "lbase", "cbase",
                                         "gebase", "gnbase", "sbase", "nlive", "nnodes", "ngroups", "pad")]


class SmInfo(C.Structure):
    """mirror of cnt_sm_info (include/cntnet_b200.h)"""
    _fields_ = [(n, C.c_int32) for n in ("done", "error", "nrounds", "max_deg", "R", "nE0", "nslots", "nelim", "nlist",
                                         "ncontrib", "nge", "ngn", "dense_nnet", "dense_nmax", "dense_nnodes",
                                         "dense_nE")] + [("dense_gpool", C.c_int64), ("reserved", C.c_int64),
                                                         ("rounds", SmRound * MAX_ROUNDS)]


ERR_SLOTS, ERR_CONTRIB, ERR_ITEMS, ERR_HASH, ERR_SCAN, ERR_ROUNDS = 1, 2, 4, 8, 16, 32


def scratch(pool, name, nbytes, device, take=False):
    """uint8 device buffer of at least nbytes.  With a pool (dict owned by an Engine) the buffer is persistent and
    grow-only with 15 % headroom: the multi-GB work areas of the direct solver differ by a fraction of a percent from
    batch to batch, and every first crossing of an allocator size class cost a 40-125 ms cudaMalloc / page mapping.
    take=True removes the buffer from the pool (the caller owns it and may hand it back with pool[name] = buf)."""
    nbytes = int(nbytes)
    if pool is None:
        return torch.empty(nbytes, dtype=torch.uint8, device=device)
    t = pool.get(name)
    if t is None or t.numel() < nbytes or t.device != torch.device(device):
        pool[name] = t = None
        t = torch.empty(nbytes + nbytes // 7 + (1 << 20), dtype=torch.uint8, device=device)
        pool[name] = t
    if take:
        pool[name] = None
    return t


class Plan(object):
    """symbolic factorisation of a batch pattern: one device buffer + the host copy of its header"""

    def __init__(self):
        self.buf = None
        self._pool = None
        self.info = SmInfo()
        self.R = self.B = self.NNZ = 0
        self.cap_slots = self.cap_contrib = 0

    # sizes the engine and the bench report
    nrounds = property(lambda s: int(s.info.nrounds))
    nnzL = property(lambda s: int(s.info.nlist))
    npairs = property(lambda s: int(s.info.ncontrib))
    nE0 = property(lambda s: int(s.info.nE0))
    nslots = property(lambda s: int(s.info.nslots))
    max_deg = property(lambda s: max(int(s.info.max_deg), int(s.info.dense_nmax)))
    dense_nmax = property(lambda s: int(s.info.dense_nmax))

    def offsets(self):
        off = (C.c_int64 * (len(ARRAYS) + 1))()
        _lib.load().cnt_sm_plan_layout(self.R, self.B, self.NNZ, self.cap_slots, self.cap_contrib, off)
        return list(off)

    def array(self, name, count=None):
        """view of one plan array (int32; doff / depos int64) -- for tests and debugging"""
        off = self.offsets()
        k = ARRAYS.index(name)
        raw = self.buf[off[k]:off[k + 1]]
        t = raw.view(torch.int64) if name in ("doff", "depos") else raw.view(torch.int32)
        return t if count is None else t[:count]

    def rounds(self):
        return [self.info.rounds[k] for k in range(self.nrounds)]

    def __del__(self):
        # hand the plan buffer back to the engine's pool (kept only if the slot is free)
        try:
            if self._pool is not None and self.buf is not None and self._pool.get("sm_plan") is None:
                self._pool["sm_plan"] = self.buf
        except Exception:
            pass


def symbolic(lib, stream_ptr, device, B, roff, R, NNZ, rowptr, col, dense_max=0, max_fill=None, rounds_hint=0,
             slot_factor=2.0, fill_factor=4.0, slack=4096, hint=None, pool=None):
    """cnt_sm_symbolic with growing capacities.  roff [B+1], rowptr [R+1], col [NNZ]: int32 device
    tensors (network-local ascending columns); hint: Plan of an earlier, similar batch (sizes the grids).
    Returns a Plan, or None when the elimination needs
    more than max_fill star-mesh contributions per round-0 edge (+100000): dense, mesh-like networks
    far above the percolation threshold, for which the caller takes an iterative solver."""
    assert roff.dtype == torch.int32 and rowptr.dtype == torch.int32 and col.dtype == torch.int32
    nE0 = NNZ // 2
    budget = None if max_fill is None else int(max_fill * max(nE0, 1) + 100000)
    plan = Plan()
    plan.R, plan.B, plan.NNZ = int(R), int(B), int(NNZ)
    vp = C.c_void_p
    plan.attempts = 0
    while True:
        plan.attempts += 1
        cap_slots = int(slot_factor * nE0) + slack
        cap_contrib = int(fill_factor * nE0) + slack
        if budget is not None:
            cap_contrib = min(cap_contrib, budget)
        cap_items = cap_contrib // 2 + cap_slots // 2 + slack
        cap_items = min(cap_items, cap_contrib + cap_slots)
        if max(cap_slots, cap_contrib, cap_items) >= (1 << 29):
            return None
        plan.cap_slots, plan.cap_contrib = cap_slots, cap_contrib
        nbytes = lib.cnt_sm_plan_layout(plan.R, plan.B, plan.NNZ, cap_slots, cap_contrib, None)
        if plan.buf is not None and pool is not None and pool.get("sm_plan") is None:
            pool["sm_plan"] = plan.buf                      # retry with larger capacities: the old buffer goes back first
        plan.buf = scratch(pool, "sm_plan", nbytes, device, take=True)
        plan._pool = pool
        ws = scratch(pool, "sm_symbolic_ws", lib.cnt_sm_symbolic_workspace_bytes(plan.R, plan.B, plan.NNZ, cap_slots,
                                                                               cap_contrib, cap_items), device)
        rc = lib.cnt_sm_symbolic(plan.B, vp(roff.data_ptr()), plan.R, plan.NNZ, vp(rowptr.data_ptr()),
                                 vp(col.data_ptr()), int(dense_max), int(rounds_hint), cap_slots, cap_contrib, cap_items,
                                 C.byref(hint.info) if hint is not None else None, vp(plan.buf.data_ptr()), plan.buf.numel(),
                                 vp(ws.data_ptr()), ws.numel(),
                                 C.byref(plan.info), stream_ptr)
        del ws
        if rc == 0:
            return plan
        err = int(plan.info.error)
        if rc != 3 or err & (ERR_SCAN | ERR_ROUNDS) or not err:
            check(rc)
        grown = False
        if err & (ERR_SLOTS | ERR_HASH):
            slot_factor *= 2
            grown = True
        if err & ERR_CONTRIB:
            if budget is not None and cap_contrib >= budget:
                return None                     # the fill budget is exhausted
            fill_factor *= 3
            grown = True
        if err & ERR_ITEMS:
            if cap_items >= cap_contrib + cap_slots:
                return None
            fill_factor *= 2
            slot_factor *= 1.5
            grown = True
        if not grown:
            check(rc)


def solve_cuda(plan, lib, stream_ptr, rowptr, val, rhs, src_eid=None, drn_eid=None, g=None, coupling=None, leak=True,
               pool=None):
    """Numeric phase on the GPU.  rowptr [R+1] int32; val [NS, >=NNZ] (CSR off-diagonals, -g); rhs [NS, >=R];
    electrode coupling of row r = g[q, src_eid[r]] + g[q, drn_eid[r]] (+ coupling[q, r]).  Returns x [NS, R]."""
    NS, R = val.shape[0], plan.R
    dev = val.device
    assert val.stride(1) == 1 and rhs.stride(1) == 1
    x = torch.empty((NS, R), dtype=torch.float64, device=dev)
    ws = scratch(pool, "sm_solve_ws", lib.cnt_sm_solve_workspace_bytes(NS, C.byref(plan.info)), dev)
    vp = C.c_void_p
    ptr = lambda t: None if t is None else vp(t.data_ptr())
    if g is not None:
        assert g.stride(1) == 1
    if coupling is not None:
        assert coupling.is_contiguous() and coupling.shape == (NS, R)
    check(lib.cnt_sm_solve(NS, plan.B, R, plan.NNZ, C.byref(plan.info), vp(plan.buf.data_ptr()), plan.cap_slots,
                           plan.cap_contrib, vp(rowptr.data_ptr()), vp(val.data_ptr()), val.stride(0), ptr(src_eid),
                           ptr(drn_eid), ptr(g), g.stride(0) if g is not None else 0, ptr(coupling), vp(rhs.data_ptr()),
                           rhs.stride(0), 1 if leak else 0, vp(x.data_ptr()), vp(ws.data_ptr()), ws.numel(), stream_ptr))
    return x


def algorithmic_bytes(plan, NS):
    """bytes the direct stage has to move at least once.  Symbolic: per round the live-edge list read twice
    (mark, incident: slot + two endpoints = 12 B) and the unknowns in play once (12 B), per incident entry
    16 B (neighbour + slot written, sorted, transposed), per contribution 12 B of plan + one 16-B hash
    entry; numeric, per configuration: every slot value written once, per incident entry two value reads
    (pivot sum, back-substitution) + its 8 B of indices twice, per contribution three value reads + 12 B
    of indices, per group one read-modify-write, 6 vectors of R doubles."""
    i = plan.info
    live = sum(int(i.rounds[k].nlive) for k in range(plan.nrounds))
    nodes = sum(int(i.rounds[k].nnodes) for k in range(plan.nrounds))
    sym = 24 * live + 12 * nodes + 16 * int(i.nlist) + 28 * int(i.ncontrib) + 16 * int(i.nslots)
    num = NS * (8 * int(i.nslots) + 32 * int(i.nlist) + 36 * int(i.ncontrib) + 16 * (int(i.nge) + int(i.ngn)) + 48 * plan.R)
    num += NS * 8 * 2 * int(i.dense_gpool)
    return sym, num
