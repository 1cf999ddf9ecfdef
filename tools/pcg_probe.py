#!/usr/bin/env python3
"""Per-iteration cost of the PCG solvers on a 100x100 um batch (fixed iteration count)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from networksim_cntfet_b200 import _lib
from networksim_cntfet_b200.engine import Engine
from networksim_cntfet_b200.elements import LinExpTransistor, conductance_table

nets = int(os.environ.get("PROBE_NETS", "8")); maxit = int(os.environ.get("PROBE_MAXIT", "2000"))
n = int(os.environ.get("PROBE_N", "100000")); scaling = float(os.environ.get("PROBE_SCALING", "100"))
methods = os.environ.get("PROBE_METHODS", "cluster,flat").split(",")
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
for method in methods:
    for rep in range(2):
        eng.stage_times_ms()
        sol = eng.solve(b, sysm, rtol=1e-30, maxit=maxit, method=method, slab_priority=[0, 3, 2, 1])
        t = eng.stage_times_ms()["pcg"]
        it = sol["iters"].cpu().numpy()
        tot = it.sum()
        byt = float((it * (12.0 * nnz + 116.0 * rows)[None, :]).sum())
        print("%-8s rep%d: %.1f ms for %d system-iterations -> %.3f us/system-iteration, %.0f GB/s algorithmic (C=%s)"
              % (method, rep, t, tot, 1e3 * t / tot, byt / t / 1e6, os.environ.get("CNT_CLUSTER_SIZE", "auto")), flush=True)
