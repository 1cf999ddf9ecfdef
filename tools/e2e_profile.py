#!/usr/bin/env python3
"""cProfile of the host side of the public ensemble call (measure_perc.measure_ensemble) on the headline workload."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import networksim_cntfet_b200 as pkg
from networksim_cntfet_b200 import measure_perc
from networksim_cntfet_b200.engine import Engine, network_seeds
eng = pkg.set_engine(Engine())
B = int(os.environ.get("E2E_BATCH", "224"))
ns = [100000] * B
for k in range(3):
    measure_perc.measure_ensemble(ns, 100, network_seeds(1, k * B, B))
torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(4):
    eng.measure_ensemble(ns, 100, network_seeds(2, k * B, B))
torch.cuda.synchronize()
t1 = time.perf_counter()
for k in range(4):
    measure_perc.measure_ensemble(ns, 100, network_seeds(3, k * B, B))
torch.cuda.synchronize()
t2 = time.perf_counter()
print("engine call %.2f ms/batch, public call %.2f ms/batch" % ((t1 - t0) / 4 * 1e3, (t2 - t1) / 4 * 1e3))
pr = cProfile.Profile()
pr.enable()
for k in range(4):
    measure_perc.measure_ensemble(ns, 100, network_seeds(4, k * B, B))
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
