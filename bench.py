#!/usr/bin/env python3
"""Headline benchmark: networks/sec on the 100x100 um ensemble (BASELINE.json configs[2]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the whole hot path over one batch of synthetic networks per GPU:
Philox stick placement -> cell-list junction finder -> union-find labels -> MNA assembly ->
the four measure_fullnet solves (Vg=0, back 10, total 10, partial 10; measure_perc.py:166-186)
by the star-mesh direct solver (--solver cluster / flat: batched two-level PCG) -> currents.  Every step draws fresh seeds.  One process per GPU; for
N>1 launch through torch.distributed.run -- ranks share nothing but a final all-gather of
the per-network statistics rows (weak scaling).

Prints ONE JSON line (rank 0).  `value` = networks/s timed with CUDA events on the engine's
stream (max over ranks); `e2e` = the same through the public host API
(networksim_cntfet_b200.measure_perc.measure_ensemble: pinned host ns/seeds in, pandas rows
out, wall clock between synchronisations); `roofline` = the solve stage's algorithmic bytes
over its CUDA-event time against the measured HBM peak; `cpu_baseline` = the numpy/scipy
oracle port of the reference path on the host cores, on a bounded sample.

`--impl reference` times that CPU port with all host cores and prints the same line with
"impl": "reference" (the unmodified reference needs >10 min per 100x100 um network and does
not exist on the GPU box; see DESIGN.md "CPU baseline").
"""
import argparse
import json
import multiprocessing
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "networks/sec at 100x100um ensemble"
UNIT = "networks/s"


def workload(args):
    return {
        "workload": "ensemble of %dx%d um networks, n=%d sticks (+2 electrodes) each, density %.3g /um^2, "
                    "measure_fullnet (4 solves: Vg=0, back/total/partial gate 10 V), LinExp element, onoffmap %d"
                    % (args.scaling, args.scaling, args.n, args.n / args.scaling ** 2, args.onoffmap),
        "networks_per_gpu_per_step": args.batch * max(1, args.engines), "batches_in_flight_per_gpu": max(1, args.engines),
        "n": args.n, "scaling": args.scaling, "onoffmap": args.onoffmap,
        "pcg_rtol": args.rtol, "solver": args.solver,
        "l2": "working set per step (%d networks x ~34 MB) exceeds the 126 MB L2; fresh seeds every step"
              % (args.batch * max(1, args.engines)),
    }


# ------------------------------------------------------------------------- CPU port (baseline)
def _cpu_one(job):
    n, scaling, seed, onoffmap = job
    from oracle import cntnet_oracle as O
    s = O.sticks_with_ends(O.canonical_sort(O.make_sticks_mt19937(n, "exp", 0.135, scaling, seed)))
    m = O.measure_fullnet(s, onoffmap=onoffmap)
    return m["ion"], m["ntests"]


def cpu_run(n, scaling, onoffmap, count, cores, seed0):
    jobs = [(n, scaling, seed0 + k, onoffmap) for k in range(count)]
    t0 = time.perf_counter()
    with multiprocessing.get_context("fork").Pool(cores) as pool:
        out = pool.map(_cpu_one, jobs, chunksize=1)
    dt = time.perf_counter() - t0
    return count / dt, dt, out


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    per_step = min(cores, args.cpu_sample or cores)
    for k in range(args.warmup):
        cpu_run(args.n, args.scaling, args.onoffmap, per_step, cores, 10_000 + k * per_step)
    t0 = time.perf_counter()
    for k in range(args.steps):
        cpu_run(args.n, args.scaling, args.onoffmap, per_step, cores, 20_000 + k * per_step)
    dt = time.perf_counter() - t0
    value = args.steps * per_step / dt
    sample = ("%d networks per step on %d processes; numpy/scipy oracle port of measure_fullnet "
              "(MT19937 sticks, cKDTree pruning + vectorised predicate, SuperLU x4)" % (per_step, cores))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------- clocks
class ClockSampler(object):
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        if not self.rows:           # timed region shorter than the polling interval: one sample right after it
            try:
                r = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS,
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10)
                self.rows = [l.strip() for l in r.stdout.splitlines() if l.strip()]
            except Exception:
                pass
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


# ------------------------------------------------------------------------- GPU arm
PCG_BYTES_PER_NNZ = 12     # val f64 + col i32 per off-diagonal entry
PCG_BYTES_PER_ROW = 116    # SpMV 28 (rowptr, diag, p, q) + update 56 (x rw, r rw, p, q, diag) + p-update 32


def pcg_algorithmic_bytes(res):
    """sum over systems of iterations x (12*nnz_offdiag + 116*rows); DESIGN.md 'Kernels'"""
    it = res["iters"].astype("float64")                     # [4, B]
    per_iter = PCG_BYTES_PER_NNZ * res["nnz_offdiag"].astype("float64") + PCG_BYTES_PER_ROW * res["rows"].astype("float64")
    return float((it * per_iter[None, :]).sum()), float(it.sum())


# ------------------------------------------------------------------------- further legs of the same JSON line
def spmv_leg(eng, side, reps=20):
    """CG SpMV against the HBM peak (the metric's third clause): y = diag .* x + A x on the reduced MNA matrices of
    ONE side x side um network (prune + series reductions, as the PCG solvers see them), the four
    measure_fullnet gate configurations per launch, through the C-ABI entry point cnt_spmv_csr (the kernel the
    row-sharded PCG runs, csrc/spmv_stream.cuh).  Working set far beyond L2; CUDA events on the engine stream."""
    import ctypes as C
    import torch
    from networksim_cntfet_b200 import _lib
    from networksim_cntfet_b200.elements import LinExpTransistor, conductance_table
    lib = _lib.load()
    old = (eng.prune, eng.series)
    eng.prune, eng.series = True, True
    try:
        n = int(10 * side * side)
        b = eng.prepare(eng.generate([n], side, seed=77))
        gates = [g for _, g in eng.fullnet_gates(b)]
        sysm = eng.assemble_values(b, gates, [conductance_table(LinExpTransistor, 1, g.levels) for g in gates])
    finally:
        eng.prune, eng.series = old
    NS, R, NNZ = len(gates), b.R, b.NNZ
    if not R:
        return {"error": "network does not percolate"}
    vp = lambda t: C.c_void_p(t.data_ptr())
    with eng._work():
        x = torch.rand((NS, R), dtype=torch.float64, device=eng.device)
        y = torch.empty((NS, R), dtype=torch.float64, device=eng.device)
        val, diag = sysm["val"], sysm["diag"]

        def launch():
            _lib.check(lib.cnt_spmv_csr(NS, R, NNZ, vp(b.rowptr), vp(b.col), vp(val), val.stride(0), vp(diag), diag.stride(0),
                                        vp(x), x.stride(0), vp(y), y.stride(0), eng.stream_ptr))
        for _ in range(3):
            launch()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(eng.stream)
        for _ in range(reps):
            launch()
        e1.record(eng.stream)
    eng.sync()
    sec = e0.elapsed_time(e1) / 1e3 / reps
    alg = NS * (12.0 * NNZ + 20.0 * R)
    out = {"kernel": "k_spmv_stream (cnt_spmv_csr)", "workload": "%gx%g um network, %d sticks: %d rows, %d off-diagonal entries, "
           "%d gate configurations per launch" % (side, side, n, R, NNZ, NS), "rows": R, "nnz_offdiag": NNZ, "configurations": NS,
           "us_per_launch": sec * 1e6, "algorithmic_bytes_per_launch": alg,
           "accounting": "SURVEY 8(d): 12 B per entry (val 8 + col 4) + 20 B per row (rowptr 4, x 8, y 8) per configuration; the "
                         "kernel also reads the 8-byte diagonal, which is not counted",
           "achieved": alg / sec / 1e9, "unit": "GB/s"}
    return out


def extra_legs(args, eng, world, rank, dist, peak):
    """driver-visible numbers for the other BASELINE configs with the CURRENT default solver, the PCG path, the SpMV"""
    import tempfile
    import numpy as np
    import torch
    from networksim_cntfet_b200 import measure_perc, netsim
    out = {}

    def sync():
        eng.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def leg(name, fn, every_rank=False):
        try:
            if every_rank or rank == 0:
                out[name] = fn()
        except Exception as e:       # a failed leg must not take the headline with it
            out[name] = {"error": "%s: %s" % (type(e).__name__, e)}
        sync()

    # config 2: ensemble of 20x20 um networks swept across the percolation threshold (measure_perc.measure_async)
    def config2():
        number = args.ens_number
        step = (8000 - 400) / max(number - 1, 1)
        cwd = os.getcwd()
        os.chdir(tempfile.mkdtemp(prefix="cntbench"))
        try:
            import contextlib, io
            with contextlib.redirect_stdout(io.StringIO()):
                # warm call of the same shape (other seeds): buffer sizes and plan capacities of this ensemble exist
                measure_perc.measure_async(1, 400, step, number, 20, onoffmap=[1], seeds=list(range(5001, 5001 + number)))
                sync()
                t0 = time.perf_counter()
                df = measure_perc.measure_async(1, 400, step, number, 20, onoffmap=[1], seeds=list(range(1, number + 1)))
                sync()
            dt = time.perf_counter() - t0
        finally:
            os.chdir(cwd)
        return {"workload": "%d networks of 20x20 um, n = 400..8000 (density 1..20 /um^2), measure_async -> rows; sharded over "
                            "%d rank(s), rows all-gathered" % (number, world), "networks": number, "seconds": dt,
                "value": number / dt, "unit": "networks/s", "timing": "wall clock between synchronisations (host API, CSV rows out)",
                "percolating_rows": int((df.ion > 0).sum()), "solver": eng.last_solver}
    leg("config2_ensemble_20um", config2, every_rank=True)

    # config 4: gate sweep of one 60x60 um device, 11 voltages x {back, partial, total} (slurm/gatesweep.sh)
    def config4():
        side, vgs = 60.0, np.linspace(-10, 10, 11)
        res = []
        for rep in range(5):
            sync_local()
            t0 = time.perf_counter()
            dev = netsim.RandomCNTNetwork(n=int(10 * side * side), scaling=side, seed=1234 + rep, onoffmap=0, rng='philox')
            sync_local()
            t1 = time.perf_counter()
            cur = dev.gate_sweep(vgs, ['back', 'partial', 'total'])
            sync_local()
            t2 = time.perf_counter()
            res.append((t1 - t0, t2 - t1, len(cur), bool(dev.percolating)))
        # wall clock of a 10-50 ms host-driven job right after the multi-GB main loops: single repetitions were seen anywhere
        # between 10 and 250 ms (allocator housekeeping of the first large requests after the ensemble batches are freed), so
        # the line quotes the fastest of repetitions 2..5 and lists all of them
        build, sweep, steps, perc = min(res[1:], key=lambda r: r[1])
        return {"workload": "60x60 um device (36000 sticks), gate sweep 11 voltages x back/partial/total = %d re-solves over one "
                            "pattern and one symbolic plan" % steps, "device_build_ms": build * 1e3, "sweep_ms": sweep * 1e3,
                "ms_per_gate_step": sweep * 1e3 / max(steps, 1), "gate_steps_per_s": steps / sweep, "percolating": perc,
                "sweep_ms_of_every_repetition": [round(r[1] * 1e3, 2) for r in res],
                "solver": eng.last_solver, "timing": "wall clock, fastest of repetitions 2..5 (rank 0)"}

    def sync_local():
        eng.sync()
        torch.cuda.synchronize()
    leg("config4_gate_sweep_60um", config4)

    # config 5 (scaled): ONE large network; N = 1: default solver on one GPU; N > 1: unknown rows sharded over the ranks
    def config5():
        side = float(args.big_side)
        n = int(10 * side * side)
        best = None
        for rep in range(2):
            sync()
            t0 = time.perf_counter()
            if world > 1:
                df, res = measure_perc.measure_fullnet_distributed(n, side, seed=5, mode="slab")
            else:
                b = eng.generate([n], side, seeds=[5])
                res = eng.measure_fullnet(b, onoffmap=1)
            sync()
            best = time.perf_counter() - t0
        return {"workload": "ONE %gx%g um network, %d sticks, measure_fullnet (4 solves)" % (side, side, n), "seconds": best,
                "sticks_per_s": n / best, "ranks": world, "solver": eng.last_solver,
                "mode": "junction search sharded by spatial slab over the ranks, junction rows exchanged over NCCL; labels, "
                        "assembly and direct solve replicated" if world > 1 else "one GPU", "ion": float(res["ion"][0]), "iters": [int(v) for v in res["iters"][:, 0]],
                "timing": "wall clock between synchronisations, second repetition"}
    leg("config5_single_network", config5, every_rank=True)

    # the iterative path (persistent cluster PCG) on a small batch of the headline workload, so that it stays measured
    def cluster():
        old = eng.solver
        eng.solver = "cluster"
        try:
            nb = 14
            from networksim_cntfet_b200.engine import network_seeds
            eng.measure_ensemble([args.n] * nb, args.scaling, network_seeds(args.seed + 7, 0, nb), onoffmap=args.onoffmap)
            eng.stage_times_ms(reset=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(eng.stream)
            res = eng.measure_ensemble([args.n] * nb, args.scaling, network_seeds(args.seed + 8, 0, nb), onoffmap=args.onoffmap)
            e1.record(eng.stream)
            eng.sync()
            st = eng.stage_times_ms(reset=True)
            nbytes, iters = pcg_algorithmic_bytes(res)
            sec = st.get("pcg", 0.0) / 1e3
            return {"workload": "%d networks of the headline workload, solver 'cluster' (two-level Jacobi PCG, persistent "
                                "thread-block-cluster kernel)" % nb, "value": nb / (e0.elapsed_time(e1) / 1e3), "unit": "networks/s",
                    "solver": eng.last_solver, "pcg_iterations": iters, "pcg_seconds": sec,
                    "equivalent_GBps": nbytes / sec / 1e9 if sec > 0 else None,
                    "note": "equivalent-work rate of an unfused HBM-streaming PCG (12 nnz + 116 rows per iteration); the kernel "
                            "works out of shared memory", "status_nonzero": int((res["status"] != 0).sum())}
        finally:
            eng.solver = old
    leg("pcg_cluster_path", cluster)

    def spmv():
        d = spmv_leg(eng, float(args.spmv_side))
        if "achieved" in d:
            d["peak"] = peak
            d["frac"] = d["achieved"] / peak
            try:
                cap = json.load(open(os.path.join(ROOT, "profiles", "r02_spmv_ncu.json")))
                d["traffic"] = cap.get("dram_bytes_total")
                d["traffic_note"] = cap.get("command")
            except Exception:
                d["traffic"] = None
        return d
    leg("spmv", spmv)
    return out


class RowGather:
    """The path's only collective: all-gather of the per-network statistics rows of a step.  gather(rows) returns the rows
    of all ranks (rank order).  gather(rows, wait=False) -- the device-timed loop -- issues the collective and returns: the
    next step starts at once, finish() waits for everything issued so far (before the timed region ends) and returns the
    rows of the last step; so the ranks, independent shards of the ensemble, do not stop for each other after every step.
    One rank: nothing to gather.  (CPU test with gloo, world 2: tests/test_distributed_cpu.py.)"""

    def __init__(self, dist, world, device):
        self.dist, self.world, self.device = dist, world, device
        self.pending = []                # (work handle, output tensors, input tensor) of the collectives not yet waited for

    def gather(self, rows, wait=True):
        if self.world == 1:
            return rows
        import numpy as np
        import torch
        t = torch.as_tensor(np.ascontiguousarray(rows), dtype=torch.float64, device=self.device)
        out = [torch.empty_like(t) for _ in range(self.world)]
        if not wait:
            self.pending.append((self.dist.all_gather(out, t, async_op=True), out, t))
            return None
        self.dist.all_gather(out, t)
        return torch.cat(out).cpu().numpy()

    def finish(self):
        import torch
        last = None
        for work, out, t in self.pending:
            work.wait()
            last = out
        self.pending = []
        return torch.cat(last).cpu().numpy() if last is not None else None


def gpu_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from networksim_cntfet_b200 import _lib, measure_perc
    from networksim_cntfet_b200.engine import Engine, network_seeds
    eng = Engine("cuda:%d" % local, profile=True)
    eng.solver = args.solver
    eng.initial_guess = args.guess
    measure_perc.set_engine(eng)
    lib = _lib.load()
    B = args.batch
    ns = [args.n] * B

    E = max(1, args.engines)               # batches of a step in flight at once (one engine = stream + host thread each)
    measure_perc.PIPELINE = E > 1

    def seeds_for(phase, step, sub=0, count=B):
        return network_seeds(args.seed + phase, ((step * E + sub) * world + rank) * B, count)

    def barrier():
        eng.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    rg = RowGather(dist, world, eng.device)
    gather_rows, finish_gathers = rg.gather, rg.finish

    def one_batch(e, sub_phase_step):
        sub, phase, step = sub_phase_step
        return e.measure_ensemble(ns, args.scaling, seeds_for(phase, step, sub), onoffmap=args.onoffmap, rtol=args.rtol,
                                  check_every=args.check_every)

    def device_step(phase, step):
        """one step = E batches of B networks; E > 1: the batches run on E engines (own stream + host thread each), so
        the latency-bound rounds of one batch's solve overlap the front end of the other"""
        jobs = [(sub, phase, step) for sub in range(E)]
        outs = eng.run_pipelined(jobs, one_batch) if E > 1 else [one_batch(eng, jobs[0])]
        rows = np.concatenate([np.stack([res["percolating"].astype(np.float64), res["ion"], res["ioff_back"],
                                         res["ioff_total"], res["ioff_partial"], res["nclust"].astype(np.float64),
                                         res["maxclust"].astype(np.float64)], 1) for res in outs])
        gather_rows(rows, wait=False)
        return outs

    # ---- clock sampler: started BEFORE the warm-up (the start-up of nvidia-smi -- NVML initialisation in a second
    # process -- stalled the first timed step by up to 35 ms), its samples counted from the start of the timed region
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- warm-up
    for k in range(args.warmup):
        device_step(0, k)
    finish_gathers()
    # everything imported and allocated so far leaves the garbage collector's generations: a full collection walks
    # ~1M module-level objects (torch, pandas, scipy) and showed up as sporadic 20-40 ms pauses inside a step
    import gc
    gc.collect()
    gc.freeze()
    eng.stage_times_ms(reset=True)
    barrier()

    # ---- timed: device-side value
    if rank == 0:
        sampler.rows = []                 # keep only the samples of the timed region
    launches0 = lib.cnt_launch_count()
    mem0 = torch.cuda.memory_stats()
    def direct_total():
        tw = getattr(eng, "_twin", None)
        return getattr(eng, "direct_bytes_total", 0) + (getattr(tw, "direct_bytes_total", 0) if tw is not None else 0)
    direct_bytes0 = direct_total()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    cur_stream = torch.cuda.current_stream()         # every engine stream is ordered after / before it (Engine._work)
    e0.record(cur_stream)
    results, step_wall = [], []
    for k in range(args.steps):
        ts = time.perf_counter()
        results.extend(device_step(1, k))
        step_wall.append(1e3 * (time.perf_counter() - ts))
    finish_gathers()                                 # every step's rows have arrived on every rank
    e1.record(cur_stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = lib.cnt_launch_count() - launches0
    mem1 = torch.cuda.memory_stats()
    direct_bytes = direct_total() - direct_bytes0
    solver_used = eng.last_solver
    direct_info = dict(getattr(eng, "last_direct", None) or {})
    dev_ms = e0.elapsed_time(e1)
    stages = eng.stage_times_ms(reset=True)
    if E > 1 and getattr(eng, "_twin", None) is not None:       # stage times of both engines (they overlap in time)
        for k, v in eng._twin.stage_times_ms(reset=True).items():
            stages[k] = stages.get(k, 0.0) + v
    clocks = sampler.stop() if rank == 0 else None
    stages_two_streams = None
    roof_steps = 0
    if E > 1:
        # Two streams: the CUDA-event time of a stage on one stream includes the kernels the other stream ran in between, so
        # the roofline objects take their stage durations (and their bytes / test counts) from a second timed region: a few
        # steps of the SAME batches on ONE engine / stream, right after the device-timed loop.  `value` is untouched.
        stages_two_streams = {k: round(v, 3) for k, v in stages.items()}
        roof_steps = min(3, args.steps)
        one_batch(eng, (0, 4, 0))                                # same shapes: nothing to warm up but the hint of the grids
        eng.stage_times_ms(reset=True)
        direct_bytes0 = direct_total()
        results = [one_batch(eng, (0, 5, k)) for k in range(roof_steps)]
        eng.sync()
        direct_bytes = direct_total() - direct_bytes0
        direct_info = dict(getattr(eng, "last_direct", None) or {})
        stages = eng.stage_times_ms(reset=True)
    roof_batches = roof_steps if E > 1 else args.steps          # batches behind `stages`, `results`, `direct_bytes`

    # ---- timed: end to end through the public host API
    for k in range(2):          # warm: pinned staging buffer, and both engines see one more pair of batches (grow-only work areas)
        measure_perc.measure_ensemble(ns * E, args.scaling, seeds_for(2, k, count=B * E), onoffmap=args.onoffmap, rtol=args.rtol)
    barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    e2e_parts = [0.0, 0.0, 0.0]
    e2e_steps = []
    for k in range(args.steps):
        ta = time.perf_counter()
        sd = seeds_for(3, k, count=B * E)
        tb = time.perf_counter()
        df, io = measure_perc.measure_ensemble(ns * E, args.scaling, sd, onoffmap=args.onoffmap,
                                               rtol=args.rtol, return_io=True)
        tc = time.perf_counter()
        fixed = np.full((3 * B * E, 2), np.nan)                      # <= 3 rows per network; equal shape on every rank
        fixed[:len(df)] = df[["ion", "ioff"]].to_numpy(dtype=np.float64)
        gather_rows(fixed)
        h2d, d2h = io["h2d_bytes"], io["d2h_bytes"]
        td = time.perf_counter()
        e2e_parts = [e2e_parts[0] + tb - ta, e2e_parts[1] + tc - tb, e2e_parts[2] + td - tc]
        e2e_steps.append(round(1e3 * (td - ta), 2))
    barrier()
    e2e_s = time.perf_counter() - t0
    mem2 = torch.cuda.memory_stats()
    memkeys = ("num_device_alloc", "num_device_free", "num_alloc_retries")
    allocator = {"timed_device_loop": {k: mem1.get(k, 0) - mem0.get(k, 0) for k in memkeys},
                 "e2e_loop": {k: mem2.get(k, 0) - mem1.get(k, 0) for k in memkeys},
                 "reserved_GB": mem2.get("reserved_bytes.all.peak", 0) / 1e9}

    # ---- reduce over ranks (max time)
    tt = torch.tensor([dev_ms, e2e_s * 1e3, t_wall * 1e3], dtype=torch.float64, device=eng.device)
    cnt = torch.tensor([float(launches)], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    dev_ms, e2e_ms, wall_ms = [float(v) for v in tt.cpu()]
    total_networks = args.steps * B * E * world
    peaks0 = {}
    try:
        peaks0 = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    extras = {} if args.no_extra else extra_legs(args, eng, world, rank, dist, float(peaks0.get("hbm_gbs", 6650.0)))
    value = total_networks / (dev_ms / 1e3)
    e2e_value = total_networks / (e2e_ms / 1e3)

    if rank == 0:
        # roofline of the dominant stage (PCG) from this rank's timed steps
        nbytes = iters = 0.0
        ntests = 0
        for res in results:
            b_, i_ = pcg_algorithmic_bytes(res)
            nbytes += b_
            iters += i_
            ntests += int(res["ntests"].sum())
        pcg_s = stages.get("pcg", 0.0) / 1e3
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = nbytes / pcg_s / 1e9 if pcg_s > 0 else 0.0
        # DRAM bytes of one launch of the dominant kernel from the committed `ncu --set full` capture of
        # this very command (profiles/): the cluster kernel keeps vectors and matrix slices on chip
        traffic = traffic_note = None
        try:
            cap = json.load(open(os.path.join(ROOT, "profiles", "r01_pcg_cluster_series_bench_launch_ncu.json")))
            if (solver_used or "").startswith("cluster") and args.batch == 56 and world == 1:
                traffic = cap["dram_bytes_total"]
                traffic_note = "ncu dram__bytes_read+write of one launch (%s)" % cap["command"]
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": "batched Jacobi-PCG (%s)" % (solver_used or args.solver), "achieved": achieved,
                    "peak": peak, "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                    "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note,
                    "algorithmic_bytes_per_launch": nbytes / max(args.steps, 1),
                    "note": "achieved = algorithmic bytes of an unfused HBM-streaming Jacobi-PCG iteration x iterations "
                            "/ CUDA-event time of the solve stage; the persistent cluster kernel serves them from "
                            "shared memory / registers (DRAM traffic ~0), so this is an equivalent-work rate",
                    "algorithmic_bytes": nbytes, "pcg_iterations": iters, "pcg_seconds": pcg_s,
                    "bytes_per_iteration": "12*nnz_offdiag + 116*rows per system"}
        if solver_used == "direct":
            # default path: star-mesh direct solve, symbolic (cnt_sm_symbolic) + numeric (cnt_sm_solve) phases, both
            # hand-written: rounds of small data-dependent passes over the live edges / pivots / neighbour pairs
            d_s = (stages.get("direct_symbolic", 0.0) + stages.get("direct_numeric", 0.0)) / 1e3
            achieved = direct_bytes / d_s / 1e9 if d_s > 0 else 0.0
            traffic = traffic_note = None
            try:
                cap = json.load(open(os.path.join(ROOT, "profiles", "r02_direct_stage_ncu.json")))
                if args.batch == cap.get("batch") and world == 1:
                    traffic = cap["dram_bytes_total"]
                    traffic_note = cap.get("note")
            except Exception:
                pass
            roofline = {"bound": "hbm", "kernel": "star-mesh direct solve (cnt_sm_symbolic: k_mark/k_select/k_incident/k_pivots/"
                                                  "k_items_a/b/scan/c per round; cnt_sm_solve: k_sm_pivot/update/back per round + "
                                                  "k_sm_dense_smem)",
                        "achieved": achieved, "peak": peak,
                        "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback", "unit": "GB/s",
                        "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note,
                        "algorithmic_bytes_per_launch": direct_bytes / max(roof_batches, 1),
                        "algorithmic_bytes": direct_bytes, "seconds": d_s, "batches": roof_batches,
                        "timed_region": ("%d steps of one batch of %d networks on ONE engine / stream right after the device-timed "
                                         "loop (under two streams the CUDA-event time of a stage includes the other stream's "
                                         "kernels)" % (roof_steps, B)) if E > 1 else "the device-timed loop",
                        "symbolic_ms_per_batch": stages.get("direct_symbolic", 0.0) / max(roof_batches, 1),
                        "numeric_ms_per_batch": stages.get("direct_numeric", 0.0) / max(roof_batches, 1),
                        "last_batch": direct_info,
                        "note": "'launch' = one batch (all rounds of both phases); algorithmic bytes per starmesh.algorithmic_bytes: "
                                "symbolic -- live-edge list read twice and unknowns in play once per round, 16 B per incident "
                                "entry, 28 B per contribution, 16 B per edge slot; numeric, per gate configuration -- 8 B per "
                                "slot, 32 B per incident entry, 36 B per contribution, 16 B per target group, 6 vectors.  The "
                                "stage is a chain of ~45 rounds of short latency-bound passes (random gathers through the "
                                "hash of stick pairs and the slot values), not a streaming kernel"}
        jt = stages.get("junctions", 0.0) / 1e3
        fp64_peak = eng.fp64_peak_tflops()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload(args), "roofline": roofline,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(cnt.item()), "clocks": clocks, "allocator_cudaMalloc_calls": allocator, "wall_ms_of_every_timed_step": [round(v, 2) for v in step_wall],
            "symbolic_attempts_total": getattr(eng, "direct_attempts", None),
            "junction_fallback_batches": getattr(eng, "junction_fallbacks", 0),
            "wall_ms_of_every_e2e_step": e2e_steps,
            "e2e_host_breakdown_ms_per_step": {"seeds": 1e3 * e2e_parts[0] / max(args.steps, 1),
                                               "measure_ensemble": 1e3 * e2e_parts[1] / max(args.steps, 1),
                                               "rows_gather": 1e3 * e2e_parts[2] / max(args.steps, 1)},
            "stages_ms_rank0": {k: round(v, 3) for k, v in stages.items()}, "stages_batches": roof_batches,
            "stages_ms_two_streams_rank0": stages_two_streams,
            "junction_tests_per_s": (ntests / jt) if jt > 0 else None, "junction_tests": ntests,
            "junction_roofline": {"bound": "fp64", "unit": "TFLOP/s", "flops_per_test": 17,
                                  "achieved": (17.0 * ntests / jt / 1e12) if jt > 0 else None,
                                  "peak": fp64_peak, "peak_source": "measured DFMA probe (cnt_fp64_peak)",
                                  "frac": (17.0 * ntests / jt / 1e12 / fp64_peak) if jt > 0 and fp64_peak else None,
                                  "note": "whole junction stage (binning + count + scan + fill); 4 of the 17 flops are "
                                          "IEEE divisions (~30 instructions each)"},
            "wall_ms_per_step": wall_ms / max(args.steps, 1),
            "percolating_fraction": float(np.mean([r["percolating"].mean() for r in results])),
            "pcg_iters_mean_per_solve": [float(np.mean([r["iters"][q].mean() for r in results])) for q in range(4)],
        }
        line["spmv"] = extras.pop("spmv", None)
        line["configs"] = extras
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            count = min(cores, args.cpu_sample or cores)
            v, dt, _ = cpu_run(args.n, args.scaling, args.onoffmap, count, cores, 30_000)
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": "%d networks of the same workload on %d processes (%.1f s); numpy/scipy oracle port of "
                          "measure_fullnet -- the unmodified reference needs ~700 s per network per core at this "
                          "size (SURVEY.md 8d) and is absent from the GPU box" % (count, cores, dt)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=160,
                    help="networks per batch (a step = engines x batch networks per GPU).  Measured, device / end to end "
                         "networks/s: 1 x 224: 3250 / 3140; 2 x 112: 3370 / 3250; 2 x 160: 3450 / 3300; 2 x 224: 3320 / 3200 "
                         "(profiles/r02_batch_sweep.txt)")
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--scaling", type=int, default=100)
    ap.add_argument("--onoffmap", type=int, default=1)
    ap.add_argument("--rtol", type=float, default=1e-15)
    ap.add_argument("--check-every", type=int, default=250)
    ap.add_argument("--solver", default="auto", choices=["auto", "cluster", "flat", "direct"])
    ap.add_argument("--seed", type=int, default=2024)
    ap.add_argument("--guess", default="zero", choices=["linear", "zero"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="networks in the CPU sample (default: one per core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--engines", type=int, default=2, choices=[1, 2],
                    help="batches of a step in flight per GPU (2: two engines = two streams + two host threads: the host work "
                         "and the latency-bound late rounds of one batch's solve overlap the kernels of the other)")
    ap.add_argument("--no-extra", action="store_true", help="skip the config 2 / 4 / 5, cluster-PCG and SpMV legs")
    ap.add_argument("--ens-number", type=int, default=1000, help="networks of the config-2 leg")
    ap.add_argument("--big-side", type=float, default=300.0, help="side (um) of the single large network of the config-5 leg")
    ap.add_argument("--spmv-side", type=float, default=600.0, help="side (um) of the network whose matrices the SpMV leg streams")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    return gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
