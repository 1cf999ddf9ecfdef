#!/usr/bin/env python3
"""BASELINE config 4: gate-voltage sweep of one device (slurm/gatesweep.sh: 60x60 um, density 9.5-12,
Vg = linspace(-10, 10, 11) x {back, partial, total} = 33 re-solves over one CSR pattern) through
netsim.RandomCNTNetwork.gate_sweep (what measure_perc.single_measure / add_voltagemeas call); every
step against the oracle (scipy direct solve + refinement) from the same sticks, which is also timed."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import networksim_cntfet_b200 as pkg
from networksim_cntfet_b200 import netsim
from networksim_cntfet_b200.engine import Engine
from oracle import cntnet_oracle as O

side = float(os.environ.get("SWEEP_SIDE", "60")); dens = float(os.environ.get("SWEEP_DENSITY", "10"))
vgnum = int(os.environ.get("SWEEP_VGNUM", "11"))
n = int(dens * side * side)
eng = pkg.set_engine(Engine(profile=True))
vgs = np.linspace(-10, 10, vgnum)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    dev = netsim.RandomCNTNetwork(n=n, scaling=side, seed=1234 + rep, onoffmap=0, rng='philox')
    torch.cuda.synchronize(); t1 = time.perf_counter()
    cur = dev.gate_sweep(vgs, ['back', 'partial', 'total'])
    torch.cuda.synchronize(); t2 = time.perf_counter()
    st = eng.stage_times_ms()
    print("rep%d %gx%g um n=%d percolating %s: device %.3f s, %d gate steps %.3f s (%.1f ms/step); solver %s; unknowns %d; status %s"
          % (rep, side, side, n, dev.percolating, t1 - t0, len(cur), t2 - t1, 1e3 * (t2 - t1) / max(len(cur), 1),
             eng.last_solver, dev._batch.R, np.unique(dev.sweep_status).tolist()), flush=True)
    print("   stages ms", {k: round(v, 1) for k, v in st.items()}, flush=True)
if os.environ.get("SWEEP_CHECK", "1") == "1" and dev.percolating:
    s = eng.sticks_to_host(dev._batch)[0]
    t0 = time.perf_counter()
    j = O.find_junctions(s)
    c = O.label_clusters(len(s["xc"]), j["stick1"], j["stick2"], s["orig"])
    t1 = time.perf_counter()
    worst, k = 0.0, 0
    for g in ['back', 'partial', 'total']:
        for vg in vgs:
            cond = O.conductance(j["kind2"], O.gate_voltages(j["x"], j["y"], g, vg), 0, "linexp")
            r = O.solve_network(len(s["xc"]), j["stick1"], j["stick2"], cond, c["label"], refine=2)
            worst = max(worst, abs(cur[k] - r["I_source"]) / r["I_source"])
            k += 1
    t2 = time.perf_counter()
    print("oracle (1 core): junctions+labels %.2f s, %d solves %.2f s; worst relative current difference %.2e"
          % (t1 - t0, k, t2 - t1, worst), flush=True)
    assert worst <= 1e-9
