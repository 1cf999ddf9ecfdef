#!/usr/bin/env python3
"""Per-iteration cost of the PCG solvers on a 100x100 um batch (fixed iteration count)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ctypes import c_void_p as C_void_p
from networksim_cntfet_b200 import _lib
from networksim_cntfet_b200.engine import Engine
from networksim_cntfet_b200.elements import LinExpTransistor, conductance_table

nets = int(os.environ.get("PROBE_NETS", "8")); maxit = int(os.environ.get("PROBE_MAXIT", "2000"))
n = int(os.environ.get("PROBE_N", "100000")); scaling = float(os.environ.get("PROBE_SCALING", "100"))
methods = os.environ.get("PROBE_METHODS", "cluster,flat").split(",")
preconds = os.environ.get("PROBE_PRECOND", "twolevel,jacobi").split(",")
eng = Engine(profile=True)
lib = _lib.load()
print("active clusters:", {c: lib.cnt_pcg_cluster_active(c) for c in (1, 2, 4, 8, 16)}, flush=True)
b = eng.generate([n] * nets, scaling, seed=1)
eng.prepare(b)
gates = [g for _, g in eng.fullnet_gates(b)]
tabs = [conductance_table(LinExpTransistor, 1, g.levels) for g in gates]
sysm = eng.assemble_values(b, gates, tabs)
rows = np.diff(np.asarray(b.h_roff)); nnz = eng.nnz_per_network(b)
print("rows", rows.tolist(), "nnz_off", nnz.tolist(), flush=True)
# aggregate setup time
with eng._work():
    sys_net = [k for q in range(4) for k in range(nets)]; sys_slab = [q for q in range(4) for k in range(nets)]
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        agg = eng._aggregates(b, sysm, sys_net, sys_slab)
        torch.cuda.synchronize(); print("aggregates setup: %.1f ms, (nagg, nseg) = %s" % (1e3 * (time.perf_counter() - t0), eng.last_aggregates), flush=True)
# A/B switches: PROBE_AB="CNT_CLUSTER_NR5" runs every case with the variable set and unset, interleaved
ab = os.environ.get("PROBE_AB", "")
variants = [(ab, v) for v in os.environ.get("PROBE_AB_VALUES", "1").split(",")] + [(ab, None)] if ab else [("", None)]
for precond in preconds:
    for method in methods:
        for rep in range(3 * len(variants)):
            var, val = variants[rep % len(variants)]
            if var:
                os.environ.pop(var, None)
                if val is not None:
                    os.environ[var] = val
                print("[%s=%s] " % (var, val), end="")
            eng.stage_times_ms()
            torch.cuda.synchronize(); t0 = time.perf_counter()
            sol = eng.solve(b, sysm, rtol=1e-30, maxit=maxit, method=method, slab_priority=[0, 3, 2, 1], precond=precond)
            torch.cuda.synchronize(); wall = 1e3 * (time.perf_counter() - t0)
            t = eng.stage_times_ms()["pcg"]
            it = sol["iters"].cpu().numpy()
            tot = it.sum()
            byt = float((it * (12.0 * nnz + 116.0 * rows)[None, :]).sum())
            print("%-8s %-8s rep%d: %.1f ms (wall %.1f) for %d system-iterations -> %.3f us/system-iteration, %.0f GB/s algorithmic"
                  % (precond, method, rep, t, wall, tot, 1e3 * t / tot, byt / t / 1e6), flush=True)

# ---- per-phase cycle breakdown of the cluster kernel (thread 0 of every CTA)
if "cluster" in methods:
    names = ["idle/queue", "setup+init", "halo refresh", "SpMV", "reduce1(+wait)", "x,r update", "segment sums",
             "barrier(seg)", "coarse+z", "reduce2(+wait)", "p update+publish", "barrier(p)", "  wait warps(r)", "  coarse values",
             "  wait warps(c)", "  segmented reduction"]
    dbg = torch.zeros(148 * 16, dtype=torch.int64, device=eng.device)
    lib.cnt_pcg_cluster_set_phase_buffer(C_void_p(dbg.data_ptr()))
    sol = eng.solve(b, sysm, rtol=1e-30, maxit=maxit, method="cluster", slab_priority=[0, 3, 2, 1], precond=preconds[0])
    torch.cuda.synchronize()
    lib.cnt_pcg_cluster_set_phase_buffer(None)
    d = dbg.cpu().numpy().reshape(148, 16)
    d = d[d.sum(1) > 0]
    its = sol["iters"].cpu().numpy().sum() / max(len(d) // 16, 1)   # iterations per cluster (approx)
    print("phase breakdown (%s), %d CTAs, ~%.0f iterations per cluster; cycles per iteration: mean / max over CTAs"
          % (preconds[0], len(d), its))
    for i, nm in enumerate(names):
        print("  %-18s %8.0f %8.0f" % (nm, d[:, i].mean() / its, d[:, i].max() / its))
    print("  %-18s %8.0f" % ("total", d.sum(1).mean() / its))
    # work phases by CTA rank (mean over the clusters): systematic imbalance shows up here
    full = dbg.cpu().numpy().reshape(148, 16)[:len(d)]
    nclu = len(d) // 16
    by_rank = full.reshape(nclu, 16, 16).mean(0) / its
    print("  cycles per iteration by CTA rank:  " + " ".join("%6s" % nm[:6] for nm in ("halo", "SpMV", "x,r", "segsum", "coarse", "p upd")))
    for r in range(16):
        print("    rank %2d                        " % r + " ".join("%6.0f" % by_rank[r, i] for i in (2, 3, 5, 6, 8, 10)))
