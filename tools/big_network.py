#!/usr/bin/env python3
"""One very large network on one GPU (towards BASELINE config 5): multi-launch two-level PCG."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from networksim_cntfet_b200.engine import Engine
from oracle import cntnet_oracle as O
side = float(os.environ.get("BIG_SIDE", "300")); dens = float(os.environ.get("BIG_DENSITY", "10"))
n = int(dens * side * side)
eng = Engine(profile=True)
for rep in range(int(os.environ.get("BIG_REPS", "2"))):
    eng.stage_times_ms()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    b = eng.generate([n], side, seed=5)
    res = eng.measure_fullnet(b, onoffmap=1)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("%gx%g um, n=%d: %.2f s total; solver %s; rows %d; junctions %d; tests %d; percolating %s; iters %s; status %s"
      % (side, side, n, dt, eng.last_solver, b.R, res["njunctions"][0], res["ntests"][0], res["percolating"][0],
         res["iters"][:, 0].tolist(), res["status"][:, 0].tolist()), flush=True)
print("stages ms", {k: round(v, 1) for k, v in eng.stage_times_ms().items()}, "ion %.12e ioff_back %.12e" % (res["ion"][0], res["ioff_back"][0]), flush=True)
if os.environ.get("BIG_CHECK", "1") == "1":
    t0 = time.perf_counter()
    s = eng.sticks_to_host(b)[0]
    m = O.measure_fullnet(s, onoffmap=1, refine=2)
    print("oracle: %.1f s; junctions %d tests %d" % (time.perf_counter() - t0, m["njunctions"], m["ntests"]))
    for k in ("ion", "ioff_back", "ioff_total", "ioff_partial"):
        print("  %-13s gpu %.12e oracle %.12e rel %.2e" % (k, res[k][0], m[k], abs(res[k][0] - m[k]) / m[k]))
    assert m["njunctions"] == res["njunctions"][0] and m["ntests"] == res["ntests"][0]
