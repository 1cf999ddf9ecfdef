"""Batched GPU pipeline: the Python side of the C-ABI (device memory = PyTorch tensors).

One `NetworkBatch` holds B independent stick networks concatenated in HBM (SoA, see
DESIGN.md "Data layout").  `Engine` drives the stages

    sticks -> junctions -> cluster labels -> CSR pattern -> (gate values -> PCG -> currents)*

each of which is one or a few calls into libcntnet_b200.so.  The reference computes the same
stages inside `RandomConductingNetwork.__init__` (netsim.py:44-47) and
`ConductionNetwork.update` (cnet.py:150-156), one network per process.
"""
import contextlib
import os
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check

AREA_PARTIAL = [0.5, 0, 0.16, 0.667]      # netsim.py:274, measure_perc.py:184
AREA_TOTAL = [0.217, 0.5, 0.167, 1.2]     # netsim.py:276, measure_perc.py:180


def network_seeds(seed, net_id0, count):
    """Per-network 64-bit seeds derived from (ensemble seed, network id) with splitmix64, so a
    network's sticks depend only on its id -- not on batch composition or GPU sharding."""
    mask = (1 << 64) - 1
    out = np.empty(count, dtype=np.uint64)
    for k in range(count):
        z = ((int(seed) & mask) * 0x9E3779B97F4A7C15 + (int(net_id0) + k + 1) * 0xBF58476D1CE4E5B9) & mask
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & mask
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & mask
        out[k] = z ^ (z >> 31)
    return out


def _p(t):
    """raw device pointer of a contiguous tensor (None -> NULL)"""
    if t is None:
        return None
    assert t.is_contiguous()
    return C.c_void_p(t.data_ptr())


class NetworkBatch(object):
    """Device-resident batch of networks.  Attributes are filled in stage by stage."""

    def __init__(self, h_soff):
        self.h_soff = [int(v) for v in h_soff]
        self.B = len(self.h_soff) - 1
        self.S = self.h_soff[-1]
        self.c_soff = _lib.i32_array(self.h_soff)
        # stage outputs
        self.J = None
        self.R = None
        self.NNZ = None

    def n_sticks(self, b):
        return self.h_soff[b + 1] - self.h_soff[b]

    @property
    def h_nzoff(self):
        """off-diagonal entries up to every network's first row (host list, fetched on first use)"""
        if getattr(self, "_h_nzoff", None) is None:
            idx = torch.as_tensor(self.h_roff, device=self.rowptr.device, dtype=torch.long)
            self._h_nzoff = [int(v) for v in self.rowptr[idx].cpu().tolist()]
        return self._h_nzoff


class GateState(object):
    """Per-junction gate-voltage state of a batch, as a small table of distinct voltage
    levels plus one u8 level id per junction (cnet.py:159-165 semantics: set_global_gate
    overwrites every junction, set_local_gate only those inside the rectangle)."""

    def __init__(self, engine, batch):
        self.engine = engine
        self.batch = batch
        self.levels = [0.0]
        with engine._work():
            self.level = torch.zeros(max(batch.J, 1), dtype=torch.uint8, device=engine.device)

    def clone(self):
        g = GateState.__new__(GateState)
        g.engine, g.batch, g.levels = self.engine, self.batch, list(self.levels)
        with self.engine._work():
            g.level = self.level.clone()
        return g

    def _level_id(self, v):
        v = float(v)
        for k, lv in enumerate(self.levels):
            if lv == v:
                return k
        if len(self.levels) >= 255:
            raise ValueError("more than 255 distinct gate voltages in one gate state")
        self.levels.append(v)
        return len(self.levels) - 1

    def set_global(self, voltage):
        self.levels = [float(voltage)]
        lib = _lib.load()
        with self.engine._work():
            check(lib.cnt_gate_levels(self.batch.J, _p(self.batch.jx), _p(self.batch.jy), _p(self.level), 1, 0,
                                      0, 0.0, 0.0, 0.0, 0.0, 0, self.engine.stream_ptr))
        return self

    def set_local(self, area, voltage):
        k = self._level_id(voltage)
        lib = _lib.load()
        with self.engine._work():
            check(lib.cnt_gate_levels(self.batch.J, _p(self.batch.jx), _p(self.batch.jy), _p(self.level), 0, 0,
                                      1, float(area[0]), float(area[1]), float(area[2]), float(area[3]), k,
                                      self.engine.stream_ptr))
        return self


class Engine(object):
    """Owns the CUDA stream and workspaces of one GPU (one process per GPU)."""

    def __init__(self, device=None, profile=False):
        if not torch.cuda.is_available():
            raise _lib.CntError("networksim_cntfet_b200 needs a CUDA device (there is no CPU fallback)")
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        torch.cuda.set_device(self.device)
        # the buffers of a batch are sized by its junction / row / entry counts, which differ by a fraction of a
        # percent from batch to batch: without rounding almost every large request misses the caching allocator's free
        # blocks and goes to cudaMalloc (a device synchronisation each).  16 size classes per power of two; expandable
        # segments, so that a request one class larger than anything cached maps a few more pages instead of a new
        # multi-GB block (measured: a 125 ms hiccup whenever a batch crossed a class boundary for the first time).
        if os.environ.get("CNT_ALLOC_ROUNDING", "1") != "0":
            try:
                setter = getattr(torch._C, "_accelerator_setAllocatorSettings", None) or torch.cuda.memory._set_allocator_settings
                setter(os.environ.get("CNT_ALLOC_CONF", "roundup_power2_divisions:16,expandable_segments:True"))
            except Exception:
                pass
        self.stream = torch.cuda.Stream(device=self.device)
        self.stream_ptr = C.c_void_p(self.stream.cuda_stream)
        self.profile = bool(profile)
        self._events = []
        self.solver = "auto"           # 'auto' | 'cluster' | 'flat'  (see Engine.solve)
        self.initial_guess = "zero"    # 'zero' | 'linear' (linear potential drop along x: measured to need ~8 % MORE
                                       # iterations, the stopping rule being relative to ||b||; kept as an option)
        self.precond = "twolevel"      # 'twolevel' (Jacobi + aggregate-level Jacobi) | 'jacobi'
        self.theta = 0.1               # strong-coupling threshold of the aggregates (relative to max |val|)
        self.cluster_compressed = True  # cluster kernel keeps its matrix slice in shared memory (class codes)
        self.last_solver = None
        self.spatial_order = True      # Z-order the unknowns of every network (cnt_assemble_reorder)
        self.prune = "auto"            # peel dangling sticks off the system (cnt_prune_leaves; exact): True | False |
                                       # "auto" = only when an iterative solver is expected (the direct solver eliminates
                                       # leaves and series sticks in its first rounds at no extra cost; measured:
                                       # tools/prune_sweep.py)
        self.slab_graph = os.environ.get("CNT_SLAB_GRAPH", "1") != "0"   # solve_sharded: replay blocks of iterations
                                       # (kernels + collectives) as one CUDA graph
        self.direct_auto = True        # solver 'auto' takes the direct solver when the pattern is sparse enough
        self.direct_max_degree = 16.0  # ... i.e. at most this many off-diagonal entries per unknown on average
        self.direct_max_fill = 40.0    # ... and the elimination creates at most this many contributions per edge
                                       # (measured: the 20x20 um density sweep 1..20 /um^2 of BASELINE config 2 needs 12.4
                                       # and runs 4.6x faster through the direct solver than through the cluster PCG)
        self.direct_dense_max = 128    # ... and finishes with one dense elimination per network once it has <= this many
                                       # unknowns left (0 = sparse rounds to the end; <= 224: the dense matrix stays in
                                       # shared memory; 128 = 69 KB, three CTAs per SM: measured 42.8 / 42.9 / 44.5 ms
                                       # symbolic + numeric per 224 networks at 128 / 160 / 192)
        self.direct_leak = True        # direct solver: answer the assembled FP64 matrix (rounded diagonal, like the
                                       # iterative solvers and the reference); False = the exact-arithmetic system
        self.series = "auto"           # eliminate an independent set of series sticks (cnt_series_select; exact;
                                       # needs prune): True | False | "auto" (as above)
        self._pool = {}                # persistent grow-only work areas of the direct solver (starmesh.scratch)
        self._direct_rounds_hint = 40  # rounds launched before the symbolic phase reads its header back
        self._direct_hint = None       # header of the previous plan (grid sizes of the round kernels)

    def twin(self):
        """A second Engine on the same GPU (own stream and events, same settings).  Two engines
        driven by two host threads pipeline an ensemble: while the persistent PCG kernel of one
        batch holds 112 SMs, the generation / junction / assembly kernels of the next batch run on
        the SMs it leaves free and the host work of both overlaps (run_pipelined)."""
        other = Engine(self.device, profile=self.profile)
        for k in ("solver", "initial_guess", "precond", "theta", "cluster_compressed", "spatial_order", "prune", "series", "slab_graph", "direct_leak", "direct_auto",
                  "direct_max_degree", "direct_max_fill", "direct_dense_max"):
            setattr(other, k, getattr(self, k))
        return other

    # ------------------------------------------------------------------ helpers
    @contextlib.contextmanager
    def _work(self, stage=None):
        """Run a stage on the engine's stream with stream-ordered semantics towards the caller:
        the engine stream first waits for the caller's current stream, and on exit the caller's
        stream waits for the engine stream, so results can be consumed with ordinary torch ops.
        (A dedicated non-default stream is needed because the PCG loop is captured into a CUDA
        graph, which the legacy default stream does not allow.)"""
        cur = torch.cuda.current_stream(self.device)
        if self.profile and stage:
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
        if cur == self.stream:
            if self.profile and stage:
                e0.record(self.stream)
            yield
            if self.profile and stage:
                e1.record(self.stream)
                self._events.append((stage, e0, e1))
            return
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            if self.profile and stage:
                e0.record(self.stream)
            yield
            if self.profile and stage:
                e1.record(self.stream)
                self._events.append((stage, e0, e1))
        cur.wait_stream(self.stream)

    def stage_times_ms(self, reset=True):
        """CUDA-event time per named stage since the last reset (profile=True only)."""
        self.stream.synchronize()
        out = {}
        for stage, e0, e1 in self._events:
            out[stage] = out.get(stage, 0.0) + e0.elapsed_time(e1)
        if reset:
            self._events = []
        return out

    def _ws(self, nbytes):
        return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=self.device)

    def _dev(self, arr, dtype):
        t = torch.as_tensor(np.ascontiguousarray(arr), dtype=dtype)
        return t.to(self.device, non_blocking=False)

    def sync(self):
        self.stream.synchronize()

    def fp64_peak_tflops(self, reps=3):
        """measured FP64 FMA peak of this GPU (roofline denominator of the junction kernel)"""
        with self._work():
            scratch = torch.empty(148 * 16 * 256, dtype=torch.float64, device=self.device)
            out = C.c_double(0.0)
            check(self.lib.cnt_fp64_peak(_p(scratch), int(reps), C.byref(out), self.stream_ptr))
        return out.value

    # ------------------------------------------------------------------ stage 0: sticks
    def batch_from_host(self, tables):
        """tables: list of dicts with xa,ya,xb,yb,xc,yc,length,kind[,orig] in POST-sort order
        (the reference's `sticks` DataFrame after netsim.py:133-134)."""
        soff = np.zeros(len(tables) + 1, dtype=np.int64)
        soff[1:] = np.cumsum([len(t["xc"]) for t in tables])
        b = NetworkBatch(soff)
        cat = lambda k, dt: np.concatenate([np.asarray(t[k], dtype=dt) for t in tables]) if tables else np.zeros(0, dt)
        with self._work():
            b.soff = self._dev(soff, torch.int32)
            for k in ("xa", "ya", "xb", "yb", "xc", "yc", "length"):
                setattr(b, k, self._dev(cat(k, np.float64), torch.float64))
            b.kind = self._dev(cat("kind", np.uint8), torch.uint8)
            if all("orig" in t for t in tables):
                b.orig = self._dev(cat("orig", np.int64), torch.int32)
            else:
                b.orig = None
        return b

    def generate(self, ns, scaling, seed=None, net_id0=0, pm=0.135, l="exp", seeds=None, seeds_dev=None):
        """netsim.py:105-134 on the GPU: len(ns) networks, network k has ns[k] body sticks plus
        the two electrodes, drawn from Philox keyed by the network's own 64-bit seed, sorted
        canonically (source, drain, then length descending).  `seeds` gives one seed per
        network (what measure_async hands out, measure_perc.py:226-233); otherwise they are
        derived from (seed, net_id0 + k); `seeds_dev` is the same array already on the device
        (int64 bit patterns).  `scaling` is a number or one per network."""
        ns = [int(n) for n in ns]
        if seeds is None and seeds_dev is not None:
            seeds = np.zeros(len(ns), dtype=np.uint64)
        if seeds is None:
            seeds = network_seeds(seed, net_id0, len(ns))
        seeds = np.asarray(seeds, dtype=np.uint64)
        assert len(seeds) == len(ns)
        soff = np.zeros(len(ns) + 1, dtype=np.int64)
        soff[1:] = np.cumsum([n + 2 for n in ns])
        b = NetworkBatch(soff)
        sc = np.broadcast_to(np.asarray(scaling, dtype=np.float64), (len(ns),)).copy()
        lib = self.lib
        with self._work("generate"):
            b.soff = self._dev(soff, torch.int32)
            dsc = self._dev(sc, torch.float64)
            S = max(b.S, 1)
            for k in ("xa", "ya", "xb", "yb", "xc", "yc", "length", "angle"):
                setattr(b, k, torch.empty(S, dtype=torch.float64, device=self.device))
            b.kind = torch.empty(S, dtype=torch.uint8, device=self.device)
            b.orig = torch.empty(S, dtype=torch.int32, device=self.device)
            ws = self._ws(lib.cnt_gen_sticks_workspace_bytes(b.S, b.B))
            len_mode = 0 if isinstance(l, str) else 1
            if isinstance(l, str) and l != "exp":
                raise ValueError("invalid L value")          # netsim.py:114-115
            # seeds travel as int64 bit patterns (torch has no uint64 arithmetic; the kernel reads u64)
            b.seeds = seeds_dev.contiguous() if seeds_dev is not None else self._dev(seeds.view(np.int64), torch.int64)
            check(lib.cnt_gen_sticks(_p(b.seeds), b.B, _p(b.soff), b.c_soff, _p(dsc),
                                     float(pm), len_mode, 0.0 if len_mode == 0 else float(l),
                                     _p(b.xc), _p(b.yc), _p(b.angle), _p(b.length), _p(b.kind),
                                     _p(b.xa), _p(b.ya), _p(b.xb), _p(b.yb), _p(b.orig),
                                     _p(ws), ws.numel(), self.stream_ptr))
        b.scaling = sc
        return b

    def sticks_to_host(self, b):
        """list of per-network dicts of numpy arrays (the reference's `sticks` columns)."""
        keys = [k for k in ("xc", "yc", "angle", "length", "kind", "xa", "ya", "xb", "yb", "orig")
                if getattr(b, k, None) is not None]
        host = {k: getattr(b, k).cpu().numpy() for k in keys}
        return [{k: host[k][b.h_soff[i]:b.h_soff[i + 1]].copy() for k in keys} for i in range(b.B)]

    # ------------------------------------------------------------------ stage 1: junctions
    def find_junctions(self, b, comm=None):
        """netsim.py:135-150.  Fills b.stick1, stick2, jx, jy, kind2, joff, ntests.  comm (slab.DistComm /
        slab.LocalComm) with more than one rank: the search is sharded by spatial slab over the ranks
        (_find_junctions_slab); every rank ends up with the full, bit-identical table."""
        if comm is not None and comm.world > 1:
            return self._find_junctions_slab(b, comm)
        lib = self.lib
        with self._work("junctions"):
            wsb = lib.cnt_junction_workspace_bytes(b.B, b.c_soff)
            b._jws = self._ws(wsb)
            jcount = torch.empty(b.S + 1, dtype=torch.int32, device=self.device)
            b.ntests = torch.zeros(b.B, dtype=torch.int64, device=self.device)
            with self._work("junction_count"):
                check(lib.cnt_junction_count(b.B, _p(b.soff), b.c_soff, _p(b.xa), _p(b.ya), _p(b.xb), _p(b.yb),
                                             _p(b.xc), _p(b.yc), _p(b.length), _p(jcount), _p(b.ntests),
                                             _p(b._jws), b._jws.numel(), self.stream_ptr))
            sws = self._ws(lib.cnt_scan_workspace_bytes(b.S + 1))
            jstart = torch.empty(b.S + 1, dtype=torch.int32, device=self.device)
            check(lib.cnt_exclusive_scan_i32(_p(jcount), _p(jstart), b.S, 1, _p(sws), sws.numel(), self.stream_ptr))
            flagp = C.c_void_p()
            check(lib.cnt_junction_errflag_ptr(b.B, b.c_soff, _p(b._jws), b._jws.numel(), C.byref(flagp)))
            off = flagp.value - b._jws.data_ptr()
            # one small D2H: total junction count + flags (bit 0: table not sorted, bit 1: temporary table overflowed)
            total, flag = torch.cat([jstart[b.S:b.S + 1], b._jws[off:off + 4].view(torch.int32)]).cpu().tolist()
            b.junction_rows_recomputed = bool(flag & 2)      # some tile ran out of buffer space: its rows are recomputed
            self.junction_fallbacks = getattr(self, "junction_fallbacks", 0) + int(bool(flag & 2))
            if flag & 1:
                raise _lib.CntError("stick table is not sorted by length descending (netsim.py:133)")
            b.J = total
            b.jstart = jstart
            n = max(total, 1)
            b.stick1 = torch.empty(n, dtype=torch.int32, device=self.device)
            b.stick2 = torch.empty(n, dtype=torch.int32, device=self.device)
            b.jx = torch.empty(n, dtype=torch.float64, device=self.device)
            b.jy = torch.empty(n, dtype=torch.float64, device=self.device)
            b.kind2 = torch.empty(n, dtype=torch.uint8, device=self.device)
            b.joff = torch.empty(b.B + 1, dtype=torch.int32, device=self.device)
            with self._work("junction_fill"):
                check(lib.cnt_junction_fill(b.B, _p(b.soff), b.c_soff, _p(b.xa), _p(b.ya), _p(b.xb), _p(b.yb),
                                            _p(b.xc), _p(b.yc), _p(b.length), _p(b.kind), _p(jstart), total,
                                            _p(b.stick1), _p(b.stick2), _p(b.jx), _p(b.jy), _p(b.kind2), _p(b.joff),
                                            _p(b._jws), b._jws.numel(), self.stream_ptr))
            b._jws = None
        return b

    def _find_junctions_slab(self, b, comm):
        """One junction search over the ranks of `comm` (BASELINE config 5: a single huge network on several GPUs).
        Every rank holds the whole stick table (it is regenerated from the seed, 70 B per stick) and evaluates the
        home sticks whose centre lies in its vertical slab of the unit box -- the FP64 tests, 98 % of the stage --
        reading the halo sticks across the slab boundary from that table (cnt_junction_count_slab / _fill_slab).
        Exchange over NCCL / NVLink: one all-reduce of the per-stick junction counts and one of the (zero-padded,
        disjoint) tables, the doubles as int64 bit patterns, so the result is bit for bit the unsharded table on
        every rank."""
        from . import slab
        lib = self.lib
        W = comm.world
        i32, i64, f64, u8 = torch.int32, torch.int64, torch.float64, torch.uint8
        with self._work("junctions"):
            wsb = lib.cnt_junction_workspace_bytes(b.B, b.c_soff)
            parts = []
            for r in comm.local_ranks:
                ws = self._ws(wsb)
                jc = torch.empty(b.S + 1, dtype=i32, device=self.device)
                nt = torch.zeros(b.B, dtype=i64, device=self.device)
                with self._work("junction_count"):
                  check(lib.cnt_junction_count_slab(b.B, _p(b.soff), b.c_soff, _p(b.xa), _p(b.ya), _p(b.xb), _p(b.yb),
                                                  _p(b.xc), _p(b.yc), _p(b.length), _p(jc), _p(nt), _p(ws), ws.numel(),
                                                    int(r), W, self.stream_ptr))
                parts.append(dict(rank=r, ws=ws, jcount=jc, ntests=nt))
            with self._work("junction_exchange"):
                comm.all_reduce([p["jcount"] for p in parts])
                comm.all_reduce([p["ntests"] for p in parts])
            jcount, b.ntests = parts[0]["jcount"], parts[0]["ntests"]
            sws = self._ws(lib.cnt_scan_workspace_bytes(b.S + 1))
            jstart = torch.empty(b.S + 1, dtype=i32, device=self.device)
            check(lib.cnt_exclusive_scan_i32(_p(jcount), _p(jstart), b.S, 1, _p(sws), sws.numel(), self.stream_ptr))
            flags = []
            for p in parts:
                flagp = C.c_void_p()
                check(lib.cnt_junction_errflag_ptr(b.B, b.c_soff, _p(p["ws"]), p["ws"].numel(), C.byref(flagp)))
                off = flagp.value - p["ws"].data_ptr()
                flags.append(p["ws"][off:off + 4].view(i32))
            head = torch.cat([jstart[b.S:b.S + 1]] + flags).cpu().tolist()
            total, flag = head[0], 0
            for f in head[1:]:
                flag |= f
            b.junction_rows_recomputed = bool(flag & 2)
            if flag & 1:
                raise _lib.CntError("stick table is not sorted by length descending (netsim.py:133)")
            b.J = total
            b.jstart = jstart
            n = max(total, 1)
            for p in parts:
                p.update(stick1=torch.zeros(n, dtype=i32, device=self.device), stick2=torch.zeros(n, dtype=i32, device=self.device),
                         jx=torch.zeros(n, dtype=f64, device=self.device), jy=torch.zeros(n, dtype=f64, device=self.device),
                         kind2=torch.zeros(n, dtype=u8, device=self.device),
                         joff=torch.empty(b.B + 1, dtype=i32, device=self.device))
                with self._work("junction_fill"):
                    check(lib.cnt_junction_fill_slab(b.B, _p(b.soff), b.c_soff, _p(b.xa), _p(b.ya), _p(b.xb), _p(b.yb),
                                                     _p(b.xc), _p(b.yc), _p(b.length), _p(b.kind), _p(jstart), total,
                                                     _p(p["stick1"]), _p(p["stick2"]), _p(p["jx"]), _p(p["jy"]),
                                                     _p(p["kind2"]), _p(p["joff"]), _p(p["ws"]), p["ws"].numel(),
                                                     int(p["rank"]), W, self.stream_ptr))
            with self._work("junction_exchange"):
                for key in ("stick1", "stick2", "jx", "jy", "kind2"):
                    slab.sum_disjoint(comm, [p[key] for p in parts])
            first = parts[0]
            b.stick1, b.stick2, b.jx, b.jy, b.kind2, b.joff = (first[k] for k in ("stick1", "stick2", "jx", "jy", "kind2", "joff"))
        return b

    # ------------------------------------------------------------------ stage 2: clusters
    def label_clusters(self, b):
        """netsim.py:175-179,197-206; measure_perc.py:76-80.  Fills b.label, b.stats."""
        lib = self.lib
        with self._work("clusters"):
            ws = self._ws(lib.cnt_cluster_workspace_bytes(b.S))
            b.label = torch.empty(max(b.S, 1), dtype=torch.int32, device=self.device)
            b.stats = torch.zeros(b.B * _lib.CNT_NSTAT, dtype=torch.int32, device=self.device)
            check(lib.cnt_label_clusters(b.B, _p(b.soff), _p(b.joff), b.S, b.J, _p(b.stick1), _p(b.stick2),
                                         _p(b.orig), _p(b.label), _p(b.stats), _p(ws), ws.numel(), self.stream_ptr))
        return b

    # ------------------------------------------------------------------ stage 3: CSR pattern
    def assemble_pattern(self, b):
        """cnet.py:104-128 structure: unknown rows + CSR pattern of L_UU (no diagonal)."""
        lib = self.lib
        with self._work("pattern"):
            ws = self._ws(lib.cnt_assemble_workspace_bytes(b.S))
            b.urow = torch.empty(max(b.S, 1), dtype=torch.int32, device=self.device)
            b.roff = torch.empty(b.B + 1, dtype=torch.int32, device=self.device)
            b.alive = b.attach = None
            prune, series = self._want_reductions(b)
            if prune and b.S > 0:
                pws = self._ws(lib.cnt_prune_workspace_bytes(b.S))
                b.alive = torch.empty(b.S, dtype=torch.uint8, device=self.device)
                b.attach = torch.empty(b.S, dtype=torch.int32, device=self.device)
                rounds = C.c_int32(0)
                check(lib.cnt_prune_leaves(b.B, _p(b.soff), _p(b.joff), b.S, b.J, _p(b.stick1), _p(b.stick2),
                                           _p(b.label), _p(b.stats), _p(b.alive), _p(b.attach), C.byref(rounds),
                                           _p(pws), pws.numel(), self.stream_ptr))
                b.prune_rounds = rounds.value
            b.elim_vid = b.voff = b.va = b.vc = b.ve1 = b.ve2 = None
            b.NV = 0
            if series and b.alive is not None and b.J > 0:
                sws_ = self._ws(lib.cnt_series_workspace_bytes(b.S))
                i32 = lambda k: torch.empty(max(k, 1), dtype=torch.int32, device=self.device)
                b.elim_vid, b.voff, b.va, b.vc, b.ve1, b.ve2 = i32(b.S), i32(b.B + 1), i32(b.S), i32(b.S), i32(b.S), i32(b.S)
                nv = C.c_int64(0)
                check(lib.cnt_series_select(b.B, _p(b.soff), _p(b.joff), b.S, b.J, _p(b.stick1), _p(b.stick2),
                                            _p(b.alive), _p(b.elim_vid), _p(b.voff), _p(b.va), _p(b.vc), _p(b.ve1),
                                            _p(b.ve2), C.byref(nv), _p(sws_), sws_.numel(), self.stream_ptr))
                b.NV = int(nv.value)
                if b.NV == 0:
                    b.elim_vid = b.voff = b.va = b.vc = b.ve1 = b.ve2 = None
            ws = self._ws(lib.cnt_assemble_workspace_bytes(b.S))
            check(lib.cnt_assemble_rows(b.B, _p(b.soff), b.S, _p(b.label), _p(b.stats), _p(b.alive), _p(b.elim_vid),
                                        _p(b.urow), _p(b.roff), _p(ws), ws.numel(), self.stream_ptr))
            hdr = torch.cat([b.roff, b.stats]).cpu().numpy()            # one D2H sync: row offsets + statistics
            b.h_roff = [int(v) for v in hdr[:b.B + 1]]
            b.h_stats = hdr[b.B + 1:].reshape(b.B, _lib.CNT_NSTAT).copy()
            b.c_roff = _lib.i32_array(b.h_roff)
            b.R = b.h_roff[-1]
            R = b.R
            if self.spatial_order and R > 0 and getattr(b, "length", None) is not None:
                rws = self._ws(lib.cnt_assemble_reorder_workspace_bytes(R, b.B))
                check(lib.cnt_assemble_reorder(b.B, b.c_roff, b.S, R, _p(b.xc), _p(b.yc), _p(b.urow), _p(rws),
                                               rws.numel(), self.stream_ptr))
            rowcnt = torch.empty(R + 1, dtype=torch.int32, device=self.device)
            check(lib.cnt_assemble_pattern_count(b.B, _p(b.soff), _p(b.joff), b.J, _p(b.stick1), _p(b.stick2),
                                                 _p(b.urow), R, _p(rowcnt), b.NV, _p(b.voff), _p(b.va), _p(b.vc),
                                                 self.stream_ptr))
            sws = self._ws(lib.cnt_scan_workspace_bytes(R + 1))
            b.rowptr = torch.empty(R + 1, dtype=torch.int32, device=self.device)
            check(lib.cnt_exclusive_scan_i32(_p(rowcnt), _p(b.rowptr), R, 1, _p(sws), sws.numel(), self.stream_ptr))
            b._h_nzoff = None
            b.NNZ = int(b.rowptr[R].item())                             # D2H sync (total off-diagonal entries)
            n = max(b.NNZ, 1)
            b.col = torch.empty(n, dtype=torch.int32, device=self.device)
            b.nz_eid = torch.empty(n, dtype=torch.int32, device=self.device)
            b.src_eid = torch.empty(max(R, 1), dtype=torch.int32, device=self.device)
            b.drn_eid = torch.empty(max(R, 1), dtype=torch.int32, device=self.device)
            b.row_stick = torch.empty(max(R, 1), dtype=torch.int32, device=self.device)
            check(lib.cnt_assemble_pattern_fill(b.B, _p(b.soff), _p(b.joff), b.J, _p(b.stick1), _p(b.stick2),
                                                _p(b.urow), _p(b.roff), R, _p(b.rowptr), _p(b.col), _p(b.nz_eid),
                                                _p(b.src_eid), _p(b.drn_eid), b.NV, _p(b.voff), _p(b.va), _p(b.vc),
                                                _p(ws), ws.numel(), self.stream_ptr))
            check(lib.cnt_assemble_row_stick(b.B, _p(b.soff), b.S, _p(b.urow), _p(b.row_stick), self.stream_ptr))
        return b

    def _want_reductions(self, b):
        """(prune, series) for this batch: the "auto" settings apply the exact reductions only when the
        solve is expected to be iterative -- solver 'cluster' / 'flat', or 'auto' with a junction graph
        too dense for the direct solver (more than direct_max_degree / 2 junctions per stick)"""
        if self.prune != "auto" and self.series != "auto":
            return bool(self.prune), bool(self.series)
        # decided from what the host already knows (no read-back): junctions per stick of the batch
        direct = self.solver == "direct" or (self.solver == "auto" and self.direct_auto and
                                              2 * b.J <= self.direct_max_degree * max(b.S, 1))
        auto = not direct
        prune = auto if self.prune == "auto" else bool(self.prune)
        series = auto if self.series == "auto" else bool(self.series)
        return prune, series and prune

    def prepare(self, b, comm=None):
        """junctions -> labels -> pattern (everything that does not depend on the gate); comm: junction search
        sharded by spatial slab over the ranks (find_junctions)"""
        self.find_junctions(b, comm)
        self.label_clusters(b)
        self.assemble_pattern(b)
        return b

    # ------------------------------------------------------------------ stage 4+5: values + solve
    def assemble_values(self, b, gates, gtables, want_cls=True):
        """One value pass per gate configuration (cnet.py:99-131).  gates: list of GateState,
        gtables: list of [9, nlevels] float64 arrays.  Returns dict(g[NS,J], val[NS,NNZ],
        diag[NS,R], rhs[NS,R]) (+ cls, gneg: the class-compressed form of the cluster PCG; want_cls=False leaves
        them out -- the direct solver does not read them and they cost two random byte gathers per entry -- and
        _ensure_classes adds them if the solve falls back to the cluster kernel)."""
        lib = self.lib
        NS = len(gates)
        R, NNZ, J = b.R, b.NNZ, b.J
        with self._work("values"):
            g = torch.empty((NS, max(J, 1)), dtype=torch.float64, device=self.device)
            val = torch.empty((NS, max(NNZ, 1)), dtype=torch.float64, device=self.device)
            diag = torch.empty((NS, max(R, 1)), dtype=torch.float64, device=self.device)
            rhs = torch.empty((NS, max(R, 1)), dtype=torch.float64, device=self.device)
            # compressed values for the cluster kernel: one class byte per entry + a 256-entry
            # table of -g per configuration (possible while 9 * nlevels <= 256)
            nv = getattr(b, "NV", 0)

            def nclasses(nl):       # base classes + (with virtual series junctions) their unordered pairs
                nb = 9 * nl
                return nb + nb * (nb + 1) // 2 if nv else nb
            use_cls = want_cls and all(nclasses(len(gs.levels)) <= 256 for gs in gates)
            cls = torch.empty((NS, max(NNZ, 1)), dtype=torch.uint8, device=self.device) if use_cls else None
            gneg = np.zeros((NS, 256), dtype=np.float64)
            for q, (gs, tab) in enumerate(zip(gates, gtables)):
                tab = np.ascontiguousarray(tab, dtype=np.float64)
                assert tab.shape == (9, len(gs.levels))
                dtab = self._dev(tab, torch.float64)
                if use_cls:
                    flat = tab.ravel()
                    gneg[q, :flat.size] = -flat
                    if nv:          # pair classes in the order k_values numbers them: (c1, c2), c1 <= c2, row-major
                        i1, i2 = np.triu_indices(flat.size)
                        with np.errstate(invalid="ignore", divide="ignore"):
                            gneg[q, flat.size:flat.size + len(i1)] = -((flat[i1] * flat[i2]) / (flat[i1] + flat[i2]))
                check(lib.cnt_assemble_values(J, _p(b.kind2), _p(gs.level), _p(dtab), tab.shape[1], R, _p(b.rowptr),
                                              _p(b.nz_eid), _p(b.src_eid), _p(b.drn_eid), _p(g[q]), _p(val[q]),
                                              _p(diag[q]), _p(rhs[q]), _p(cls[q]) if use_cls else None,
                                              _p(getattr(b, "ve1", None)), _p(getattr(b, "ve2", None)),
                                              self.stream_ptr))
            out = dict(g=g, val=val, diag=diag, rhs=rhs)
            if use_cls:
                out.update(cls=cls, gneg=self._dev(np.nan_to_num(gneg, nan=0.0), torch.float64))
            elif not want_cls:
                out["_cls_args"] = (gates, gtables)
        return out

    def _ensure_classes(self, b, sysm):
        """class-compressed values for the cluster PCG when assemble_values left them out (direct solve expected, but
        the elimination exceeded its fill budget)"""
        if "cls" in sysm or "_cls_args" not in sysm:
            return
        gates, gtables = sysm.pop("_cls_args")
        full = self.assemble_values(b, gates, gtables, want_cls=True)
        for k in ("cls", "gneg"):
            if k in full:
                sysm[k] = full[k]

    def _aggregates(self, b, sysm, sys_net, sys_slab, chunk_C=0):
        """tables of the two-level preconditioner (cnt_pcg_aggregates); returns (struct, keepalive)"""
        lib = self.lib
        NS = sysm["val"].shape[0]
        n = NS * b.R
        nsys = len(sys_net)
        # segments end every 256 members and, for the cluster tables, at every CTA chunk boundary: up to one per row
        cap_seg = n + 2 if chunk_C else n // 2 + n // 256 + 2
        i32 = lambda k: torch.empty(max(k, 1), dtype=torch.int32, device=self.device)
        t = dict(agg_of_row=i32(n), mem=i32(n), seg_start=i32(cap_seg), seg_count=i32(cap_seg), seg_agg=i32(cap_seg),
                 seg_sys=i32(cap_seg), agg_seg0=i32(n // 2 + 2), sys_seg=i32(2 * nsys),
                 row_q=i32(2 * n), cta_seg_off=i32(nsys * chunk_C + 1),
                 cta_seg=i32(4 * (cap_seg + nsys * chunk_C) if chunk_C else 1),
                 cta_agg_off=i32(nsys * chunk_C + 1), cta_agg=i32(cap_seg if chunk_C else 1),
                 row_slot=i32(n if chunk_C else 1), seg_home=i32(cap_seg if chunk_C else 1),
                 row_einv=torch.empty(max(n, 1), dtype=torch.float64, device=self.device),
                 agg_einv=torch.empty(n // 2 + 2, dtype=torch.float64, device=self.device))
        ws = self._ws(lib.cnt_pcg_aggregate_workspace_bytes(nsys, NS, b.R))
        nagg, nseg = C.c_int64(0), C.c_int64(0)
        check(lib.cnt_pcg_aggregates(b.B, nsys, NS, _lib.i32_array(sys_net), _lib.i32_array(sys_slab), b.c_roff,
                                     _p(b.roff), b.R, b.NNZ, _p(b.rowptr), _p(b.col), _p(sysm["val"]), _p(sysm["diag"]),
                                     float(self.theta), _p(t["agg_of_row"]), _p(t["mem"]), _p(t["seg_start"]),
                                     _p(t["seg_count"]), _p(t["seg_agg"]), _p(t["seg_sys"]), _p(t["agg_seg0"]),
                                     _p(t["agg_einv"]), _p(t["sys_seg"]), int(chunk_C), _p(t["row_q"]), _p(t["row_einv"]),
                                     _p(t["cta_seg_off"]), _p(t["cta_seg"]), _p(t["cta_agg_off"]), _p(t["cta_agg"]),
                                     _p(t["row_slot"]), _p(t["seg_home"]), cap_seg, C.byref(nagg), C.byref(nseg), _p(ws),
                                     ws.numel(),
                                     self.stream_ptr))
        st = _lib.Aggregates(t["agg_of_row"].data_ptr(), t["mem"].data_ptr(), t["seg_start"].data_ptr(),
                             t["seg_count"].data_ptr(), t["seg_agg"].data_ptr(), t["seg_sys"].data_ptr(),
                             t["agg_seg0"].data_ptr(), t["agg_einv"].data_ptr(), t["sys_seg"].data_ptr(),
                             t["row_q"].data_ptr(), t["row_einv"].data_ptr(), t["cta_seg_off"].data_ptr(),
                             t["cta_seg"].data_ptr(), t["cta_agg_off"].data_ptr(), t["cta_agg"].data_ptr(),
                             t["row_slot"].data_ptr(), t["seg_home"].data_ptr(), nagg.value, nseg.value,
                             int(chunk_C), 0)
        self.last_aggregates = (nagg.value, nseg.value)
        return st, t

    def solve(self, b, sysm, x0=None, rtol=1e-15, maxit=2000000, check_every=250, nets=None, method=None,
              slab_priority=None, precond=None):
        """Batched Jacobi-PCG (replaces cnet.py:147-149).  sysm: output of assemble_values for
        NS gate configurations.  Every (configuration q, percolating network) pair is one
        system of a single solver call.  method: 'cluster' (persistent thread-block-cluster
        kernel, default when every system fits on chip), 'flat' (multi-launch kernels).
        slab_priority: configurations expected to need more iterations first (work-queue order).
        Returns dict(x[NS,R], iters[NS,B], relres[NS,B], status[NS,B]) (non-percolating networks:
        iters 0, status 0)."""
        lib = self.lib
        NS = sysm["val"].shape[0]
        R, NNZ, B = b.R, b.NNZ, b.B
        nets = [n for n in (range(B) if nets is None else nets) if b.h_roff[n + 1] > b.h_roff[n]]
        method = method or self.solver
        with self._work("pcg"):
            x = torch.zeros((NS, max(R, 1)), dtype=torch.float64, device=self.device) if x0 is None else x0.clone()
            iters = torch.zeros((NS, B), dtype=torch.int32, device=self.device)
            relres = torch.zeros((NS, B), dtype=torch.float64, device=self.device)
            status = torch.zeros((NS, B), dtype=torch.int32, device=self.device)
            if R > 0 and nets:
                slabs = list(range(NS))
                if slab_priority is not None:
                    slabs.sort(key=lambda q: -slab_priority[q])
                # big systems first inside a configuration: better tail behaviour of the work queue
                nets_sorted = sorted(nets, key=lambda n: -(b.h_roff[n + 1] - b.h_roff[n]))
                sys_net = [n for q in slabs for n in nets_sorted]
                sys_slab = [q for q in slabs for n in nets_sorted]
                nsys = len(sys_net)
                it = torch.zeros(nsys, dtype=torch.int32, device=self.device)
                rr = torch.zeros(nsys, dtype=torch.float64, device=self.device)
                stt = torch.zeros(nsys, dtype=torch.int32, device=self.device)
                max_rows = max(b.h_roff[n + 1] - b.h_roff[n] for n in nets)
                precond = self.precond if precond is None else precond
                auto = method == "auto"
                xd = None
                if method == "direct" or (auto and self._direct_ok(b)):
                    xd = self._direct_solve(b, sysm, max_fill=self.direct_max_fill if auto else None)
                    if xd is not None:
                        method = "direct"
                if auto and xd is None:
                    method = "cluster" if max_rows <= lib.cnt_pcg_cluster_max_rows() else "flat"
                if method == "cluster" and self.cluster_compressed:
                    self._ensure_classes(b, sysm)
                compressed = method == "cluster" and self.cluster_compressed and "cls" in sysm
                chunk_C = 0
                if compressed:
                    max_nnz = max(b.h_nzoff[n + 1] - b.h_nzoff[n] for n in nets)
                    chunk_C = lib.cnt_pcg_cluster_size(max_rows, max_nnz)
                if auto and compressed and lib.cnt_pcg_cluster_active(chunk_C) <= 0:
                    method, compressed, chunk_C = "flat", False, 0      # cluster launch not possible here
                with self._work("pcg_aggregates"):
                    agg = (self._aggregates(b, sysm, sys_net, sys_slab, chunk_C)
                           if precond == "twolevel" and method != "direct" else None)
                self.last_solver = method
                if method == "direct":
                    x = xd
                elif method == "cluster":
                    ws = self._ws(lib.cnt_pcg_cluster_workspace_bytes(nsys, NS, R))
                    with self._work("pcg_kernel"):
                      check(lib.cnt_pcg_cluster_solve(B, nsys, NS, _lib.i32_array(sys_net), _lib.i32_array(sys_slab),
                                                    b.c_roff, R, NNZ, _p(b.rowptr), _p(b.col), _p(sysm["val"]),
                                                    _p(sysm.get("cls")) if self.cluster_compressed else None,
                                                    _p(sysm.get("gneg")) if self.cluster_compressed else None,
                                                    _p(sysm["diag"]), _p(sysm["rhs"]), _p(x), float(rtol), int(maxit),
                                                    _p(it), _p(rr), _p(stt),
                                                    C.byref(agg[0]) if (agg and self.cluster_compressed and
                                                                        "cls" in sysm) else None,
                                                    chunk_C, _p(ws), ws.numel(), self.stream_ptr))
                    self._cluster_fallback(b, sysm, x, sys_net, sys_slab, it, rr, stt, rtol, maxit, check_every)
                elif method == "flat":
                    ws = self._ws(lib.cnt_pcg_workspace_bytes(nsys, NS, R))
                    check(lib.cnt_pcg_solve(B, nsys, NS, _lib.i32_array(sys_net), _lib.i32_array(sys_slab), b.c_roff,
                                            R, NNZ, _p(b.rowptr), _p(b.col), _p(sysm["val"]), _p(sysm["diag"]),
                                            _p(sysm["rhs"]), _p(x), float(rtol), int(maxit), int(check_every),
                                            _p(it), _p(rr), _p(stt), C.byref(agg[0]) if agg else None,
                                            _p(ws), ws.numel(), self.stream_ptr))
                else:
                    raise ValueError("unknown solver %r" % method)
                idx_q = torch.as_tensor(sys_slab, device=self.device, dtype=torch.long)
                idx_n = torch.as_tensor(sys_net, device=self.device, dtype=torch.long)
                iters[idx_q, idx_n] = it
                relres[idx_q, idx_n] = rr
                status[idx_q, idx_n] = stt
        return dict(x=x, iters=iters, relres=relres, status=status)

    def _direct_ok(self, b):
        """auto: is the batch pattern sparse enough for the direct solver?"""
        return bool(self.direct_auto and b.R and b.NNZ <= self.direct_max_degree * b.R)

    def _direct_solve(self, b, sysm, max_fill=None):
        """Star-mesh elimination (starmesh.py, csrc/starmesh.cu) of all systems of the batch at once:
        symbolic phase once per batch pattern (cached on the batch: gate sweeps reuse it), numeric
        phase for all configurations of sysm together.  Returns x [NS, R]."""
        from . import starmesh
        R, NNZ = b.R, b.NNZ
        plan = getattr(b, "_sm_plan", None)
        if plan is False:
            if max_fill is not None:
                return None             # an earlier attempt exceeded the fill budget
            plan = None
        if plan is None:
            with self._work("direct_symbolic"):
                plan = starmesh.symbolic(self.lib, self.stream_ptr, self.device, b.B, b.roff, R, NNZ, b.rowptr, b.col,
                                         dense_max=int(min(self.direct_dense_max, 1024)),       # k_sm_dense: <= 1400 rows
                                         max_fill=max_fill, rounds_hint=self._direct_rounds_hint,
                                         hint=self._direct_hint, pool=self._pool)
                if plan is None:
                    b._sm_plan = False
                    return None
                b._sm_plan = plan
                self.direct_attempts = getattr(self, "direct_attempts", 0) + plan.attempts    # > 1 per plan: capacities grew
                self._direct_rounds_hint = min(plan.nrounds + 4, 400)     # first read-back of the next batch
                self._direct_hint = starmesh.Plan()                        # header only: sizes the next batch's grids
                self._direct_hint.info = plan.info
        with self._work("direct_numeric"):
            x = starmesh.solve_cuda(plan, self.lib, self.stream_ptr, b.rowptr, sysm["val"], sysm["rhs"],
                                    src_eid=b.src_eid, drn_eid=b.drn_eid, g=sysm["g"], leak=self.direct_leak, pool=self._pool)
        sym, num = starmesh.algorithmic_bytes(plan, sysm["val"].shape[0])
        self.last_direct = dict(rounds=plan.nrounds, nnzL=plan.nnzL, max_deg=plan.max_deg, npairs=plan.npairs,
                                nE0=plan.nE0, nslots=plan.nslots, dense_nmax=plan.dense_nmax,
                                bytes_symbolic=sym, bytes_numeric=num)
        self.direct_bytes_total = getattr(self, "direct_bytes_total", 0) + sym + num
        return x

    def solve_sharded(self, b, sysm, comm, rtol=1e-15, maxit=2000000, check_every=50):
        """Two-level PCG of ONE network (b.B == 1) whose unknown rows are sharded over the ranks of
        `comm` (slab.DistComm: one rank per process / GPU over NCCL; slab.LocalComm(W): W ranks inside
        this process).  Every rank holds the same prepared batch and the same assembled values; it
        keeps its row range of them, and p's boundary values plus three small sums are exchanged
        per iteration (networksim_cntfet_b200/slab.py, csrc/pcg_slab.cu).  Same outputs as solve()."""
        from . import slab
        assert b.B == 1, "solve_sharded takes one network"
        NS = sysm["val"].shape[0]
        R, NNZ = b.R, b.NNZ
        with self._work("pcg"):
            iters = torch.zeros((NS, 1), dtype=torch.int32, device=self.device)
            relres = torch.zeros((NS, 1), dtype=torch.float64, device=self.device)
            status = torch.zeros((NS, 1), dtype=torch.int32, device=self.device)
            if R == 0:
                return dict(x=torch.zeros((NS, 1), dtype=torch.float64, device=self.device), iters=iters,
                            relres=relres, status=status)
            agg_of_row = agg_einv = keep = None
            if self.precond == "twolevel":
                with self._work("pcg_aggregates"):
                    _, keep = self._aggregates(b, sysm, [0] * NS, list(range(NS)), 0)
                    agg_of_row = keep["agg_of_row"][:NS * R]
                    agg_einv = keep["agg_einv"][:max(int(self.last_aggregates[0]), 1)]
            ph = slab.CudaPhases(self.stream_ptr)
            with self._work("slab_plan"):
                states = slab.build_states(b.rowptr[:R + 1], b.col[:max(NNZ, 1)], sysm["val"][:, :NNZ],
                                           sysm["diag"][:, :R], sysm["rhs"][:, :R], agg_of_row, agg_einv, comm,
                                           ph.block_rows(), rtol, maxit)
            with self._work("slab_pcg"):
                self.last_slab_info = {}
                slab.run_pcg(states, comm, ph, check_every, graph=self.slab_graph, info=self.last_slab_info)
                x = slab.gather_solution(states, comm, R)
            sc = states[0].scal
            iters[:, 0] = sc[:, _lib.CNT_SLAB_ITERS].to(torch.int32)
            status[:, 0] = sc[:, _lib.CNT_SLAB_STATUS].to(torch.int32)
            bb = sc[:, _lib.CNT_SLAB_BB]
            relres[:, 0] = torch.where(bb > 0, (sc[:, _lib.CNT_SLAB_RR] / bb.clamp(min=1e-300)).sqrt(), torch.zeros_like(bb))
            self.last_solver = "slab x%d" % comm.world
            self.last_slab_states = states
        return dict(x=x, iters=iters, relres=relres, status=status)

    def _cluster_fallback(self, b, sysm, x, sys_net, sys_slab, it, rr, stt, rtol, maxit, check_every):
        """systems the cluster kernel reported as not fitting on chip (status 3) are re-solved by
        the multi-launch kernels; costs one small D2H of the status vector"""
        lib = self.lib
        st = stt.cpu().numpy()
        bad = np.flatnonzero(st == 3)
        if len(bad) == 0:
            return
        NS = sysm["val"].shape[0]
        net2 = [sys_net[k] for k in bad]
        slab2 = [sys_slab[k] for k in bad]
        n2 = len(bad)
        it2 = torch.zeros(n2, dtype=torch.int32, device=self.device)
        rr2 = torch.zeros(n2, dtype=torch.float64, device=self.device)
        st2 = torch.zeros(n2, dtype=torch.int32, device=self.device)
        ws = self._ws(lib.cnt_pcg_workspace_bytes(n2, NS, b.R))
        check(lib.cnt_pcg_solve(b.B, n2, NS, _lib.i32_array(net2), _lib.i32_array(slab2), b.c_roff, b.R, b.NNZ,
                                _p(b.rowptr), _p(b.col), _p(sysm["val"]), _p(sysm["diag"]), _p(sysm["rhs"]), _p(x),
                                float(rtol), int(maxit), int(check_every), _p(it2), _p(rr2), _p(st2), None, _p(ws),
                                ws.numel(), self.stream_ptr))
        idx = torch.as_tensor(bad, device=self.device, dtype=torch.long)
        it[idx] = it2
        rr[idx] = rr2
        stt[idx] = st2
        self.last_solver = "cluster+flat"

    # ------------------------------------------------------------------ stage 6: write-back
    def currents(self, b, g, x):
        """cnet.py:132-146 for one configuration: V[S], |I|[J], isrc[B,3]."""
        lib = self.lib
        with self._work("currents"):
            V = torch.empty(max(b.S, 1), dtype=torch.float64, device=self.device)
            I = torch.empty(max(b.J, 1), dtype=torch.float64, device=self.device)
            isrc = torch.zeros((b.B, 3), dtype=torch.float64, device=self.device)
            ws = self._ws(lib.cnt_currents_workspace_bytes(b.B, b.J))
            check(lib.cnt_currents(b.B, _p(b.soff), _p(b.joff), b.S, b.J, _p(b.stick1), _p(b.stick2), _p(b.label),
                                   _p(b.stats), _p(b.urow), _p(getattr(b, "attach", None)),
                                   _p(getattr(b, "elim_vid", None)), _p(getattr(b, "va", None)),
                                   _p(getattr(b, "vc", None)), _p(getattr(b, "ve1", None)), _p(getattr(b, "ve2", None)),
                                   _p(g), _p(x), _p(V), _p(I), _p(isrc), _p(ws), ws.numel(), self.stream_ptr))
        return V, I, isrc

    # ------------------------------------------------------------------ measure_fullnet, batched
    def fullnet_gates(self, b):
        """The four gate configurations of measure_fullnet (measure_perc.py:166-186):
        Vg=0, back gate 10, 'total' top gate 10, 'partial' top gate 10."""
        with self._work():
            g0 = GateState(self, b)
            gb = GateState(self, b).set_global(10)
            gt = GateState(self, b).set_global(0).set_local(AREA_TOTAL, 10)
            gp = GateState(self, b).set_global(0).set_local(AREA_PARTIAL, 10)
        return [("ion", g0), ("back", gb), ("total", gt), ("partial", gp)]

    def linear_guess(self, b, NS):
        """x0[q, row] = 0.1 * (0.99 - xc) / 0.98 clipped to [0, 0.1]: the potential of a homogeneous
        sheet between the electrodes at x=0.01 (0.1 V) and x=0.99 (0 V) (netsim.py:121-124,181-182).
        Any x0 gives the same answer; a good one saves iterations."""
        with self._work():
            R = b.R
            if R == 0 or getattr(b, "length", None) is None:
                return None
            roff = torch.as_tensor(b.h_roff, device=self.device, dtype=torch.int64)
            soff = torch.as_tensor(b.h_soff, device=self.device, dtype=torch.int64)
            rows = torch.arange(R, device=self.device, dtype=torch.int64)
            net = torch.bucketize(rows, roff[1:], right=True)
            stick = b.row_stick[:R].to(torch.int64) + soff[net]
            v = (0.1 * (0.99 - b.xc[stick]) / 0.98).clamp_(0.0, 0.1)
            return v.unsqueeze(0).repeat(NS, 1).contiguous()

    def solve_gates(self, b, gates, onoffmap, element, rtol=1e-15, maxit=2000000, check_every=250, want_fields=False,
                    row_shard=None):
        """Assemble + solve + write back a list of GateStates.  Returns dict with isrc[NS,B,3],
        iters/relres/status[NS,B] (+ V[NS,S], I[NS,J], g[NS,J], x when want_fields)."""
        from .elements import conductance_table
        tabs = [conductance_table(element, onoffmap, gs.levels) for gs in gates]
        direct_expected = row_shard is None and (self.solver == "direct" or (self.solver == "auto" and self._direct_ok(b)))
        sysm = self.assemble_values(b, gates, tabs, want_cls=not direct_expected)
        prio = [max(abs(v) for v in gs.levels) for gs in gates]   # higher |Vg| -> more iterations -> first
        x0 = self.linear_guess(b, len(gates)) if self.initial_guess == "linear" else None
        if (row_shard is None and b.B == 1 and self.solver == "auto" and self.precond == "twolevel" and
                b.R > self.lib.cnt_pcg_cluster_max_rows() and not self._direct_ok(b)):
            # one network beyond the cluster kernel's on-chip capacity: the row-slab solver with a single
            # rank advances all configurations together (measured 2x the multi-launch solver at 10M sticks)
            from . import slab
            row_shard = slab.LocalComm(1)
        if row_shard is not None:      # one network, rows sharded over the ranks of a communicator
            sol = self.solve_sharded(b, sysm, row_shard, rtol=rtol, maxit=maxit, check_every=min(check_every, 50))
        else:
            sol = self.solve(b, sysm, x0=x0, rtol=rtol, maxit=maxit, check_every=check_every, slab_priority=prio)
        NS = len(gates)
        out = dict(iters=sol["iters"], relres=sol["relres"], status=sol["status"])
        isrc = []
        Vs, Is = [], []
        for q in range(NS):
            V, I, cur = self.currents(b, sysm["g"][q], sol["x"][q])
            isrc.append(cur)
            if want_fields:
                Vs.append(V)
                Is.append(I)
        with self._work():
            out["isrc"] = torch.stack(isrc)
            if want_fields:
                out.update(V=torch.stack(Vs), I=torch.stack(Is), g=sysm["g"], x=sol["x"], sysm=sysm)
        return out

    def measure_fullnet(self, b, onoffmap=1, element=None, rtol=1e-15, maxit=2000000, check_every=250,
                        want_fields=False, slab_shard=None, row_shard=None):
        """Batched body of measure_perc.measure_fullnet (measure_perc.py:140-203) for a prepared
        batch: returns a dict of host numpy arrays, one entry per network.
        slab_shard=(rank, world): solve only the gate configurations q with q % world == rank (the
        others are left zero) -- used to spread ONE large network over several GPUs: every rank
        builds the same network from the same seed and the currents are summed afterwards.
        row_shard=comm (slab.DistComm / slab.LocalComm): all configurations are solved together with
        the unknown rows sharded over the communicator's ranks (solve_sharded); every rank returns
        the full result."""
        from .elements import LinExpTransistor
        element = element or LinExpTransistor
        if b.R is None:
            self.prepare(b)
        gates = self.fullnet_gates(b)
        mine = list(range(len(gates)))
        if slab_shard is not None:
            mine = [q for q in mine if q % slab_shard[1] == slab_shard[0]]
        st = b.h_stats
        perc = st[:, _lib.STAT_PERCOLATING].astype(bool)
        ng = len(gates)
        isrc = np.zeros((ng, b.B, 3))
        iters = np.zeros((ng, b.B), dtype=np.int32)
        relres = np.zeros((ng, b.B))
        status = np.zeros((ng, b.B), dtype=np.int32)
        res = None
        if mine:
            res = self.solve_gates(b, [gates[q][1] for q in mine], onoffmap, element, rtol, maxit, check_every,
                                   want_fields, row_shard=row_shard)
            self.sync()
            isrc[mine] = res["isrc"].cpu().numpy()
            iters[mine] = res["iters"].cpu().numpy()
            relres[mine] = res["relres"].cpu().numpy()
            status[mine] = res["status"].cpu().numpy()
        cur = np.where(perc[None, :], isrc[:, :, 2], 0.0)     # power form (see currents.cu)
        # sheet conductance of the square sample, G_sq = I_source / Vds (Vds = 0.1 V, netsim.py:182), per
        # gate configuration [ion, back, total, partial]
        sheet = cur / 0.1
        out = dict(percolating=perc, sheet_conductance=sheet, nclust=st[:, _lib.STAT_NCLUST_REF].copy(),
                   maxclust=st[:, _lib.STAT_MAXCLUST_REF].copy(),
                   nclust_true=st[:, _lib.STAT_NCLUST_TRUE].copy(), maxclust_true=st[:, _lib.STAT_MAXCLUST_TRUE].copy(),
                   ncomp=st[:, _lib.STAT_NCOMP].copy(), ecomp=st[:, _lib.STAT_ECOMP].copy(),
                   ion=cur[0], ioff_back=cur[1], ioff_total=cur[2], ioff_partial=cur[3],
                   isrc_all=isrc, iters=iters, relres=relres, status=status, ntests=b.ntests.cpu().numpy(),
                   njunctions=np.diff(b.joff.cpu().numpy()))
        if want_fields:
            out["fields"] = res
        return out

    # ------------------------------------------------------------------ generic junction graphs
    def batch_from_junctions(self, n, stick1, stick2, jx, jy, kind2):
        """One network given directly as a junction list over n nodes (node 0 = source, node 1 =
        ground), e.g. a networkx graph handed to cnet.ConductionNetwork.  Rows must be sorted
        by (stick1, stick2) with stick1 < stick2."""
        b = NetworkBatch([0, int(n)])
        J = len(stick1)
        with self._work():
            b.soff = self._dev(np.array([0, n]), torch.int32)
            b.joff = self._dev(np.array([0, J]), torch.int32)
            pad = lambda a, dt: np.ascontiguousarray(a if J else np.zeros(1), dtype=dt)
            b.stick1 = self._dev(pad(stick1, np.int32), torch.int32)
            b.stick2 = self._dev(pad(stick2, np.int32), torch.int32)
            b.jx = self._dev(pad(jx, np.float64), torch.float64)
            b.jy = self._dev(pad(jy, np.float64), torch.float64)
            b.kind2 = self._dev(pad(kind2, np.uint8), torch.uint8)
            b.ntests = torch.zeros(1, dtype=torch.int64, device=self.device)
            b.xc = torch.full((max(n, 1),), float("nan"), dtype=torch.float64, device=self.device)
            b.yc = b.xc.clone()
            b.orig = None
        b.J = J
        return b

    def attach_junctions(self, b, stick1, stick2, jx, jy, kind2):
        """Use a junction table computed elsewhere (netsim.load_system reads one from CSV,
        netsim.py:238-249) for a single-network batch; rows sorted by (stick1, stick2)."""
        assert b.B == 1
        J = len(stick1)
        with self._work():
            pad = lambda a, dt: np.ascontiguousarray(a if J else np.zeros(1), dtype=dt)
            b.joff = self._dev(np.array([0, J]), torch.int32)
            b.stick1 = self._dev(pad(stick1, np.int32), torch.int32)
            b.stick2 = self._dev(pad(stick2, np.int32), torch.int32)
            b.jx = self._dev(pad(jx, np.float64), torch.float64)
            b.jy = self._dev(pad(jy, np.float64), torch.float64)
            b.kind2 = self._dev(pad(kind2, np.uint8), torch.uint8)
            b.ntests = torch.zeros(1, dtype=torch.int64, device=self.device)
        b.J = J
        return b

    def solve_given_conductances(self, b, g_host, rtol=1e-15, maxit=2000000, check_every=250):
        """update() for conductances evaluated by the caller (one configuration)."""
        lib = self.lib
        R, NNZ, J = b.R, b.NNZ, b.J
        with self._work("values"):
            g = self._dev(np.asarray(g_host, dtype=np.float64).reshape(1, -1) if J else np.zeros((1, 1)), torch.float64)
            val = torch.empty((1, max(NNZ, 1)), dtype=torch.float64, device=self.device)
            diag = torch.empty((1, max(R, 1)), dtype=torch.float64, device=self.device)
            rhs = torch.empty((1, max(R, 1)), dtype=torch.float64, device=self.device)
            check(lib.cnt_assemble_values(J, None, None, None, 1, R, _p(b.rowptr), _p(b.nz_eid), _p(b.src_eid),
                                          _p(b.drn_eid), _p(g[0]), _p(val[0]), _p(diag[0]), _p(rhs[0]), None,
                                          _p(getattr(b, "ve1", None)), _p(getattr(b, "ve2", None)), self.stream_ptr))
        sysm = dict(g=g, val=val, diag=diag, rhs=rhs)
        sol = self.solve(b, sysm, rtol=rtol, maxit=maxit, check_every=check_every)
        V, I, cur = self.currents(b, g[0], sol["x"][0])
        with self._work():
            out = dict(iters=sol["iters"], relres=sol["relres"], status=sol["status"], isrc=cur[None],
                       V=V[None], I=I[None], g=g, x=sol["x"], sysm=sysm)
        return out

    def measure_ensemble(self, ns, scaling, seeds, onoffmap=1, element=None, l="exp", pm=0.135, rtol=1e-15,
                         maxit=2000000, check_every=250, seeds_dev=None):
        """Batched body of measure_perc.measure_async (measure_perc.py:205-242): one network
        per (ns[k], seeds[k]); generation, junctions, clusters, assembly, the four
        measure_fullnet solves and the write-back all run on this GPU.  Host in: the ns/seeds
        arrays; host out: one row of statistics per network (dict of numpy arrays)."""
        b = self.generate(ns, scaling, seeds=seeds, l=l, pm=pm, seeds_dev=seeds_dev)
        out = self.measure_fullnet(b, onoffmap=onoffmap, element=element, rtol=rtol, maxit=maxit,
                                   check_every=check_every)
        out["sticks"] = np.asarray([int(n) for n in ns])
        out["seed"] = np.asarray(seeds, dtype=np.uint64)
        out["rows"] = np.diff(np.asarray(b.h_roff))
        # per-network entry counts cost one more read-back: only fetched for the iterative solvers' traffic accounting
        out["nnz_offdiag"] = self.nnz_per_network(b) if self.last_solver != "direct" else np.zeros(b.B, dtype=np.int64)
        return out

    def run_pipelined(self, jobs, fn, other=None):
        """Run fn(engine, job) for every job, alternating between this engine and its twin on two
        host threads (job k on engine k % 2), each thread inside its engine's stream so that
        nothing is ordered through the shared default stream.  Results in job order.  Networks
        are independent and seeded individually, so the results do not depend on the split."""
        import threading
        if len(jobs) < 2:
            return [fn(self, j) for j in jobs]
        if other is None:
            if getattr(self, "_twin", None) is None:
                self._twin = self.twin()
            other = self._twin
        engines = (self, other)
        cur = torch.cuda.current_stream(self.device)
        out = [None] * len(jobs)
        err = []

        def worker(w):
            eng = engines[w]
            try:
                torch.cuda.set_device(eng.device)
                eng.stream.wait_stream(cur)
                with torch.cuda.stream(eng.stream):
                    for k in range(w, len(jobs), 2):
                        out[k] = fn(eng, jobs[k])
            except BaseException as e:      # re-raised in the caller
                err.append(e)

        threads = [threading.Thread(target=worker, args=(w,)) for w in range(2)]
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        for eng in engines:
            cur.wait_stream(eng.stream)
        if err:
            raise err[0]
        return out

    def nnz_per_network(self, b):
        """off-diagonal non-zeros of every network's reduced matrix (for traffic accounting)"""
        return np.diff(np.asarray(b.h_nzoff))
