import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import networksim_cntfet_b200 as pkg
from networksim_cntfet_b200 import measure_perc
from networksim_cntfet_b200.engine import Engine, network_seeds
eng = pkg.set_engine(Engine(profile=True))
B = 224; ns = [100000] * B
def run(tag, phase, step, public):
    sd = network_seeds(2024 + phase, step * B, B)
    eng.stage_times_ms()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    if public:
        measure_perc.measure_ensemble(ns, 100, sd)
    else:
        eng.measure_ensemble(ns, 100, sd)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    st = eng.stage_times_ms()
    print(tag, phase, step, "public" if public else "engine", "%.1f ms" % (dt * 1e3), {k: round(v, 1) for k, v in st.items() if v > 3},
          getattr(eng, "last_direct", {}).get("rounds"), eng.junction_fallbacks if hasattr(eng, "junction_fallbacks") else None, flush=True)
for k in range(3): run("warm", 0, k, False)
for k in range(3): run("dev ", 1, k, False)
run("pub ", 2, 0, True)
for k in range(3): run("pub ", 3, k, True)
run("pub again", 3, 0, True)
run("eng again", 3, 0, False)
