#!/usr/bin/env python3
"""GPU experiment: do the exact reductions (leaf pruning, series elimination) still pay with the direct solver?"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from networksim_cntfet_b200.engine import Engine, network_seeds

eng = Engine("cuda:0", profile=True)
B = int(os.environ.get("BATCH", "56"))
eng.direct_dense_max = int(os.environ.get("DENSE", "192"))
for prune, series in ((True, True), (True, False), (False, False)):
    eng.prune, eng.series = prune, series
    for rep in range(4):
        if rep == 1:
            eng.stage_times_ms(reset=True)
        res = eng.measure_ensemble([100000] * B, 100, network_seeds(7, rep * B, B), onoffmap=1)
    eng.sync()
    st = eng.stage_times_ms(reset=True)
    d = eng.last_direct
    print(json.dumps(dict(prune=prune, series=series, rows=int(res["rows"].sum()), rounds=d["rounds"], nnzL=d["nnzL"],
                          npairs=d["npairs"], stages={k: round(v / 3, 2) for k, v in st.items()}, ion0=float(res["ion"][0]))))
