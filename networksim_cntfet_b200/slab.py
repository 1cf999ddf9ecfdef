"""ONE very large network sharded over several GPUs (BASELINE config 5).

The reference cannot spread one network over several workers (its unit of parallelism is the
process, measure_perc.py:230-235); this extends solve_mna (cnet.py:147-149).  Every rank builds
the same network from the same seed (counter-based RNG; sticks, junctions, labels and the CSR
pattern are identical everywhere -- the front end is < 1 % of the run time at 10M sticks), then
the SOLVE, which is > 99 % of it, is sharded: rank r owns the contiguous range
[R r / W, R (r+1) / W) of the Z-ordered unknown rows -- a compact spatial patch of the device --,
keeps that slice of the matrix and of the vectors, and the ranks exchange per iteration

  * the values of p on the patch boundaries (halo rows), one all-gather of the exported rows,
  * two small all-reduces: p.q, and [r.z partial, r.r, sums of the aggregates of the two-level
    preconditioner that span ranks]

for all gate configurations of measure_fullnet at once.  The per-rank arithmetic is CUDA
(csrc/pcg_slab.cu through the C-ABI, `CudaPhases`); this module owns the partition plan (index
arithmetic with torch ops on the device, once per network) and the exchange schedule
(torch.distributed over NCCL / NVLink, `DistComm`).  `LocalComm` runs the W ranks of a plan inside
one process on one GPU -- the same kernels, plans and schedule with the collectives replaced by
copies -- which is how the single-GPU parity tests cover the multi-rank logic.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import check


# --------------------------------------------------------------------------- communicators
class DistComm(object):
    """torch.distributed process group; one local rank."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.local_ranks = [self.rank]

    def all_reduce(self, tensors):
        self.dist.all_reduce(tensors[0], group=self.group)

    def all_gather(self, recvs, sends):
        """recvs[0]: [world, ...] contiguous, sends[0]: [...]"""
        self.dist.all_gather_into_tensor(recvs[0].view(-1), sends[0].view(-1), group=self.group)


class LocalComm(object):
    """All W ranks in this process (tests and single-GPU runs): collectives are copies."""

    def __init__(self, world):
        self.world = int(world)
        self.rank = 0
        self.local_ranks = list(range(self.world))

    def all_reduce(self, tensors):
        if self.world == 1:
            return
        s = torch.stack(list(tensors)).sum(0)
        for t in tensors:
            t.copy_(s)

    def all_gather(self, recvs, sends):
        if self.world == 1:
            recvs[0].view(-1).copy_(sends[0].view(-1))
            return
        cat = torch.stack(list(sends))
        for r in recvs:
            r.copy_(cat)


def junction_slab_owner(tile_col, tiles_per_side, world):
    """rank that evaluates the home sticks of a tile column in the slab-sharded junction search (mirror of
    slab_owner in csrc/junctions.cu): contiguous, balanced slabs of whole tile columns"""
    return (int(tile_col) * int(world)) // int(tiles_per_side)


def sum_disjoint(comm, tensors):
    """all-reduce (sum) of per-rank arrays with DISJOINT supports (zero wherever another rank wrote): floating-point
    arrays travel as their int64 bit patterns, so every element -- including -0.0 and NaN payloads -- arrives
    unchanged.  tensors: one per local rank of `comm`, reduced in place."""
    views = [t.view(torch.int64) if t.dtype == torch.float64 else t for t in tensors]
    comm.all_reduce(views)


# --------------------------------------------------------------------------- plan
class SlabState(object):
    """Slice of the system owned by one rank + work vectors (all tensors on one device)."""


def _bounds(R, world):
    return [(R * r) // world for r in range(world + 1)]


def _lookup(sorted_ids, keys):
    """position of every key in the ascending list sorted_ids, -1 where it is absent"""
    if sorted_ids.numel() == 0:
        return torch.full_like(keys, -1)
    pos = torch.searchsorted(sorted_ids, keys).clamp(max=int(sorted_ids.numel()) - 1)
    return torch.where(sorted_ids[pos] == keys, pos, torch.full_like(pos, -1))


def _pad2(rows, width, fill, dtype, device):
    out = torch.full((len(rows), max(int(width), 1)), fill, dtype=dtype, device=device)
    for q, t in enumerate(rows):
        if t.numel():
            out[q, :t.numel()] = t.to(dtype)
    return out


def build_states(rowptr, col, val, diag, rhs, agg_of_row, agg_einv, comm, block_rows, rtol, maxit):
    """Partition plan + buffers for every local rank of `comm`.
    rowptr [R+1], col [NNZ] (int32), val [NS, NNZ], diag, rhs [NS, R]: the replicated system;
    agg_of_row [NS*R] (-1 = none), agg_einv [nagg]: aggregates of the two-level preconditioner
    (cnt_pcg_aggregates), or None for plain Jacobi."""
    dev = rowptr.device
    NS, R = diag.shape[0], diag.shape[1]
    W = comm.world
    bounds = _bounds(R, W)
    bt = torch.tensor(bounds, dtype=torch.int64, device=dev)
    i32, i64, f64 = torch.int32, torch.int64, torch.float64

    # aggregates that span ranks: same list (ascending aggregate id) on every rank
    shared_ids = []
    if agg_of_row is not None:
        a_all = agg_of_row.view(NS, R).to(i64)
        nagg = int(agg_einv.numel())
        rows = torch.arange(R, device=dev, dtype=i64)
        for q in range(NS):
            a = a_all[q]
            m = a >= 0
            lo_r = torch.full((nagg,), R, dtype=i64, device=dev).scatter_reduce(0, a[m], rows[m], "amin")
            hi_r = torch.full((nagg,), -1, dtype=i64, device=dev).scatter_reduce(0, a[m], rows[m], "amax")
            has = hi_r >= 0
            span = has & (torch.bucketize(lo_r.clamp(max=R - 1), bt[1:], right=True) !=
                          torch.bucketize(hi_r.clamp(min=0), bt[1:], right=True))
            shared_ids.append(torch.nonzero(span).flatten())
    else:
        shared_ids = [torch.zeros(0, dtype=i64, device=dev) for _ in range(NS)]
    maxSh = max(1, max(int(s.numel()) for s in shared_ids))

    states = []
    for rank in comm.local_ranks:
        st = SlabState()
        lo, hi = bounds[rank], bounds[rank + 1]
        n_own = hi - lo
        k0, k1 = int(rowptr[lo]), int(rowptr[hi])
        st.rank, st.lo, st.hi, st.NS, st.n_own, st.R = rank, lo, hi, NS, n_own, R
        st.rowptr = (rowptr[lo:hi + 1] - k0).to(i32).contiguous()
        cg = col[k0:k1].to(i64)
        inside = (cg >= lo) & (cg < hi)
        st.halo_rows = torch.unique(cg[~inside])                     # ascending global rows
        st.n_halo = int(st.halo_rows.numel())
        cl = torch.where(inside, cg - lo, n_own + torch.searchsorted(st.halo_rows, cg))
        st.col = cl.to(i32).contiguous()
        st.val = val[:, k0:k1].contiguous()
        st.diag = diag[:, lo:hi].contiguous() if n_own else torch.ones((NS, 1), dtype=f64, device=dev)
        st.rhs = rhs[:, lo:hi].contiguous() if n_own else torch.zeros((NS, 1), dtype=f64, device=dev)
        # the matrix is symmetric: the rows other ranks need from me are my rows with a remote column
        rowid = torch.repeat_interleave(torch.arange(n_own, device=dev, dtype=i64), st.rowptr.to(i64).diff())
        st.exp_rows = torch.unique(rowid[~inside]).to(i32).contiguous()
        st.n_exp = int(st.exp_rows.numel())
        # vectors
        z = lambda *shape: torch.zeros(shape, dtype=f64, device=dev)
        st.x, st.r, st.q, st.dinv = z(NS, max(n_own, 1)), z(NS, max(n_own, 1)), z(NS, max(n_own, 1)), z(NS, max(n_own, 1))
        st.p = z(NS, max(n_own + st.n_halo, 1))
        st.nblk = (max(n_own, 1) + block_rows - 1) // block_rows
        st.part = z(NS, 3, st.nblk)
        st.red_pq = z(NS)
        st.red2 = z(NS, 2 + maxSh)
        st.scal = z(NS, _lib.CNT_SLAB_NSCAL)
        st.done = torch.zeros(NS, dtype=i32, device=dev)
        st.maxSh = maxSh
        st.rtol, st.maxit = float(rtol), int(maxit)
        # aggregates of the own rows
        row_agg, aggptr, aggmem, agg_sh, einv, sh_loc, einv_sh, big = [], [], [], [], [], [], [], []
        for q in range(NS):
            sid = shared_ids[q]
            if agg_of_row is None:
                own = torch.full((n_own,), -1, dtype=i64, device=dev)
            else:
                own = a_all[q, lo:hi]
            m = own >= 0
            uniq, inv = torch.unique(own[m], return_inverse=True)           # ascending aggregate ids
            ra = torch.full((n_own,), -1, dtype=i64, device=dev)
            ra[m] = inv
            row_agg.append(ra)
            order = torch.sort(inv, stable=True)[1]                          # members ascending inside an aggregate
            aggmem.append(torch.nonzero(m).flatten()[order])
            cnt = torch.bincount(inv, minlength=int(uniq.numel()))
            aggptr.append(torch.cat([torch.zeros(1, dtype=i64, device=dev), cnt.cumsum(0)]))
            slot = _lookup(sid, uniq)                                         # shared slot of a local aggregate
            hit = slot >= 0
            agg_sh.append(slot)
            einv.append(agg_einv[uniq] if agg_of_row is not None else torch.zeros(0, dtype=f64, device=dev))
            sl = torch.full((int(sid.numel()),), -1, dtype=i64, device=dev)
            sl[slot[hit]] = torch.nonzero(hit).flatten()
            sh_loc.append(sl)
            einv_sh.append(agg_einv[sid] if agg_of_row is not None else torch.zeros(0, dtype=f64, device=dev))
            big.append(torch.nonzero(cnt > _lib.CNT_SLAB_BIGAGG).flatten())
        st.nloc = torch.tensor([int(t.numel()) for t in einv], dtype=i32, device=dev)
        st.nsh = torch.tensor([int(t.numel()) for t in shared_ids], dtype=i32, device=dev)
        st.nbig = torch.tensor([int(t.numel()) for t in big], dtype=i32, device=dev)
        st.max_nbig = max(int(t.numel()) for t in big)
        st.maxL = max(1, max(int(t.numel()) for t in einv))
        st.maxBig = max(1, st.max_nbig)
        st.row_agg = _pad2(row_agg, n_own, -1, i32, dev)
        st.aggmem = _pad2(aggmem, n_own, 0, i32, dev)
        st.aggptr = torch.zeros((NS, st.maxL + 1), dtype=i32, device=dev)
        for q, t in enumerate(aggptr):
            st.aggptr[q, :t.numel()] = t.to(i32)
            st.aggptr[q, t.numel():] = int(t[-1])
        st.agg_shared = _pad2(agg_sh, st.maxL, -1, i32, dev)
        st.einv = _pad2(einv, st.maxL, 0.0, f64, dev)
        st.shared_local = _pad2(sh_loc, maxSh, -1, i32, dev)
        st.einv_sh = _pad2(einv_sh, maxSh, 0.0, f64, dev)
        st.bigagg = _pad2(big, st.maxBig, 0, i32, dev)
        st.S = z(NS, st.maxL)
        states.append(st)

    # exchange of the export lists (global rows, ascending, padded with -1)
    nexp = [torch.tensor([st.n_exp], dtype=i64, device=dev) for st in states]
    nexp_all = [torch.zeros((W, 1), dtype=i64, device=dev) for _ in states]
    comm.all_gather(nexp_all, nexp)
    sizes = [int(v) for v in nexp_all[0].flatten().tolist()]
    emax = max(1, max(sizes))
    mine = []
    for st in states:
        t = torch.full((emax,), -1, dtype=i64, device=dev)
        t[:st.n_exp] = st.exp_rows.to(i64) + st.lo
        mine.append(t)
    lists = [torch.zeros((W, emax), dtype=i64, device=dev) for _ in states]
    comm.all_gather(lists, mine)
    for st, all_lists in zip(states, lists):
        st.emax = emax
        st.sendbuf = torch.zeros((NS, emax), dtype=f64, device=dev)
        st.recvbuf = torch.zeros((W, NS, emax), dtype=f64, device=dev)
        owner = torch.bucketize(st.halo_rows, bt[1:], right=True)
        pos = torch.zeros_like(owner)
        for s in range(W):
            sel = owner == s
            if bool(sel.any()):
                lst = all_lists[s, :sizes[s]]
                ps = torch.searchsorted(lst, st.halo_rows[sel])
                if not bool((lst[ps.clamp(max=max(sizes[s] - 1, 0))] == st.halo_rows[sel]).all()):
                    raise _lib.CntError("slab plan: a halo row is missing from its owner's export list "
                                        "(matrix pattern not symmetric)")
                pos[sel] = ps
        st.halo_src = (owner * (NS * emax) + pos).to(i64).contiguous()
    return states


# --------------------------------------------------------------------------- CUDA phases
class SlabCtx(C.Structure):
    """mirror of cnt_slab_ctx (include/cntnet_b200.h)"""
    _fields_ = [("NS", C.c_int32), ("nblk", C.c_int32), ("n_own", C.c_int64), ("n_halo", C.c_int64),
                ("val_stride", C.c_int64), ("rowptr", C.c_void_p), ("col", C.c_void_p), ("val", C.c_void_p),
                ("diag", C.c_void_p), ("rhs", C.c_void_p), ("x", C.c_void_p), ("r", C.c_void_p), ("q", C.c_void_p),
                ("dinv", C.c_void_p), ("p", C.c_void_p),
                ("maxL", C.c_int64), ("maxSh", C.c_int64), ("maxBig", C.c_int64),
                ("nloc", C.c_void_p), ("nsh", C.c_void_p), ("nbig", C.c_void_p),
                ("max_nbig", C.c_int32), ("pad0", C.c_int32),
                ("aggptr", C.c_void_p), ("aggmem", C.c_void_p), ("row_agg", C.c_void_p), ("agg_shared", C.c_void_p),
                ("einv", C.c_void_p), ("shared_local", C.c_void_p), ("einv_sh", C.c_void_p), ("bigagg", C.c_void_p),
                ("S", C.c_void_p),
                ("part", C.c_void_p), ("red_pq", C.c_void_p), ("red2", C.c_void_p), ("scal", C.c_void_p),
                ("done", C.c_void_p),
                ("n_exp", C.c_int64), ("emax", C.c_int64), ("exp_rows", C.c_void_p), ("sendbuf", C.c_void_p),
                ("recvbuf", C.c_void_p), ("halo_src", C.c_void_p),
                ("rtol", C.c_double), ("maxit", C.c_int32), ("pad1", C.c_int32)]


class CudaPhases(object):
    """The per-rank phases on the GPU (csrc/pcg_slab.cu)."""

    def __init__(self, stream_ptr):
        self.lib = _lib.load()
        self.stream_ptr = stream_ptr

    @staticmethod
    def block_rows():
        return int(_lib.load().cnt_slab_block_rows())

    def ctx(self, st):
        c = getattr(st, "_ctx", None)
        if c is not None:
            return c
        ptr = lambda t: t.data_ptr() if t is not None and t.numel() else None
        for name in ("rowptr", "col", "val", "diag", "rhs", "x", "r", "q", "dinv", "p", "nloc", "nsh", "nbig", "aggptr",
                     "aggmem", "row_agg", "agg_shared", "einv", "shared_local", "einv_sh", "bigagg", "S", "part",
                     "red_pq", "red2", "scal", "done", "exp_rows", "sendbuf", "recvbuf", "halo_src"):
            t = getattr(st, name)
            assert t.is_cuda and t.is_contiguous(), name
        c = SlabCtx(NS=st.NS, nblk=st.nblk, n_own=st.n_own, n_halo=st.n_halo, val_stride=st.val.shape[1],
                    rowptr=ptr(st.rowptr), col=ptr(st.col), val=ptr(st.val), diag=ptr(st.diag), rhs=ptr(st.rhs),
                    x=ptr(st.x), r=ptr(st.r), q=ptr(st.q), dinv=ptr(st.dinv), p=ptr(st.p),
                    maxL=st.maxL, maxSh=st.maxSh, maxBig=st.maxBig, nloc=ptr(st.nloc), nsh=ptr(st.nsh),
                    nbig=ptr(st.nbig), max_nbig=st.max_nbig, pad0=0, aggptr=ptr(st.aggptr), aggmem=ptr(st.aggmem),
                    row_agg=ptr(st.row_agg), agg_shared=ptr(st.agg_shared), einv=ptr(st.einv),
                    shared_local=ptr(st.shared_local), einv_sh=ptr(st.einv_sh), bigagg=ptr(st.bigagg), S=ptr(st.S),
                    part=ptr(st.part), red_pq=ptr(st.red_pq), red2=ptr(st.red2), scal=ptr(st.scal), done=ptr(st.done),
                    n_exp=st.n_exp, emax=st.emax, exp_rows=ptr(st.exp_rows), sendbuf=ptr(st.sendbuf),
                    recvbuf=ptr(st.recvbuf), halo_src=ptr(st.halo_src), rtol=st.rtol, maxit=st.maxit, pad1=0)
        st._ctx = c
        return c

    def init(self, st):
        check(self.lib.cnt_slab_init(C.byref(self.ctx(st)), self.stream_ptr))

    def direction(self, st, first):
        check(self.lib.cnt_slab_direction(C.byref(self.ctx(st)), int(first), self.stream_ptr))

    def spmv(self, st):
        check(self.lib.cnt_slab_spmv(C.byref(self.ctx(st)), self.stream_ptr))

    def update(self, st):
        check(self.lib.cnt_slab_update(C.byref(self.ctx(st)), self.stream_ptr))


# --------------------------------------------------------------------------- schedule
def _iteration(states, comm, phases):
    for st in states:
        phases.spmv(st)
    comm.all_reduce([st.red_pq for st in states])
    for st in states:
        phases.update(st)
    comm.all_reduce([st.red2 for st in states])
    for st in states:
        phases.direction(st, 0)
    comm.all_gather([st.recvbuf for st in states], [st.sendbuf for st in states])


def run_pcg(states, comm, phases, check_every=50, graph=False, info=None):
    """Advance all local ranks in lock step until every configuration has stopped.  Returns the
    number of iterations issued.  graph=True (CUDA): the first block of `check_every` iterations
    runs eagerly (it also warms up the communicator), then one block -- kernels of all phases and
    the collectives between them -- is captured into a CUDA graph on the current stream and
    replayed; the host only polls the done flags between replays."""
    for st in states:
        phases.init(st)
    comm.all_reduce([st.red2 for st in states])
    for st in states:
        phases.direction(st, 1)
    comm.all_gather([st.recvbuf for st in states], [st.sendbuf for st in states])
    issued = 0
    maxit = states[0].maxit
    g = None
    while issued < maxit:
        if g is not None:
            g.replay()
        else:
            for _ in range(check_every):
                _iteration(states, comm, phases)
        issued += check_every
        if bool(states[0].done.all()):        # D2H sync; the flags are identical on every rank
            break
        if graph and g is None:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=torch.cuda.current_stream(), capture_error_mode="thread_local"):
                for _ in range(check_every):
                    _iteration(states, comm, phases)
    if info is not None:
        info["graph"] = g is not None
        info["issued"] = issued
    return issued


def gather_solution(states, comm, R):
    """x [NS, R] on every rank from the ranks' own slices"""
    st0 = states[0]
    W = comm.world
    width = max(b - a for a, b in zip(_bounds(R, W)[:-1], _bounds(R, W)[1:]))
    sends, recvs = [], []
    for st in states:
        s = torch.zeros((st.NS, max(width, 1)), dtype=torch.float64, device=st.x.device)
        s[:, :st.n_own] = st.x[:, :st.n_own]
        sends.append(s)
        recvs.append(torch.zeros((W, st.NS, max(width, 1)), dtype=torch.float64, device=st.x.device))
    comm.all_gather(recvs, sends)
    bounds = _bounds(R, W)
    x = torch.zeros((st0.NS, max(R, 1)), dtype=torch.float64, device=st0.x.device)
    for r in range(W):
        x[:, bounds[r]:bounds[r + 1]] = recvs[0][r, :, :bounds[r + 1] - bounds[r]]
    return x
