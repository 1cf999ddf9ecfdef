import os, sys, time
sys.path.insert(0, "/root/repo")
import torch
from networksim_cntfet_b200.engine import Engine
from networksim_cntfet_b200 import slab
eng = Engine(profile=True)
side = float(os.environ.get("BIG_SIDE", "1000")); n = int(10 * side * side)
for rep in range(3):
    for W in (1, 2, 8):
        b = eng.generate([n], side, seeds=[5])
        eng.stage_times_ms()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        eng.find_junctions(b, slab.LocalComm(W) if W > 1 else None)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print("rep", rep, "W", W, "wall %.1f ms" % (dt * 1e3), {k: round(v, 1) for k, v in eng.stage_times_ms().items()}, flush=True)
