#!/usr/bin/env python3
"""BASELINE config 2: ensemble of 20x20 um networks swept across the percolation threshold
(density 1..20 sticks/um^2) through the public measure_perc.measure_async entry point."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import networksim_cntfet_b200 as pkg
from networksim_cntfet_b200 import measure_perc
from networksim_cntfet_b200.engine import Engine

number = int(os.environ.get("ENS_NUMBER", "1000"))
eng = pkg.set_engine(Engine(profile=True))
eng.solver = os.environ.get("ENS_SOLVER", "auto")
if os.environ.get("ENS_MAX_FILL"):
    eng.direct_max_fill = float(os.environ["ENS_MAX_FILL"])
if os.environ.get("ENS_MAX_DEGREE"):
    eng.direct_max_degree = float(os.environ["ENS_MAX_DEGREE"])
os.makedirs("/tmp/ens20", exist_ok=True); os.chdir("/tmp/ens20")
seeds = list(range(1, number + 1))
step = (8000 - 400) / max(number - 1, 1)
for rep in range(2):
    eng.stage_times_ms()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    df = measure_perc.measure_async(1, 400, step, number, 20, onoffmap=[1], seeds=seeds)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    st = eng.stage_times_ms()
    nets = df.seed.nunique()
    print("rep%d: %d networks (n = 400..8000) in %.2f s -> %.1f networks/s; percolating rows %d; stages ms %s"
          % (rep, nets, dt, nets / dt, int((df.ion > 0).sum()), {k: round(v, 1) for k, v in st.items()}), flush=True)
    print("   solver", eng.last_solver, getattr(eng, "last_direct", None), flush=True)
