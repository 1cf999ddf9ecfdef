#!/usr/bin/env python3
"""GPU experiment: stage times of the direct solver against Engine.direct_dense_max (batch of 56 networks, 100x100 um)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from networksim_cntfet_b200.engine import Engine, network_seeds

eng = Engine("cuda:0", profile=True)
B = int(os.environ.get("BATCH", "56"))
for dm in [int(v) for v in os.environ.get("DENSE", "64,128,192,256,384,512,768").split(",")]:
    eng.direct_dense_max = dm
    for rep in range(3):
        res = eng.measure_ensemble([100000] * B, 100, network_seeds(7, rep * B, B), onoffmap=1)
        eng.sync()
        if rep == 0:
            eng.stage_times_ms(reset=True)
    st = eng.stage_times_ms(reset=True)
    d = eng.last_direct
    print(json.dumps(dict(dense_max=dm, rounds=d["rounds"], nnzL=d["nnzL"], npairs=d["npairs"], nslots=d["nslots"],
                          dense_nmax=d["dense_nmax"], symbolic_ms=round(st["direct_symbolic"] / 2, 3),
                          numeric_ms=round(st["direct_numeric"] / 2, 3), ion0=float(res["ion"][0]))))
