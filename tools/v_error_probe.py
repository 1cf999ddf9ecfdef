#!/usr/bin/env python3
"""Voltage error of the GPU solve against the refined oracle on full-size networks (per gate)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from networksim_cntfet_b200.engine import Engine
from oracle import cntnet_oracle as O

eng = Engine()
eng.prune = os.environ.get("VPRUNE", "1") == "1"
eng.solver = os.environ.get("VSOLVER", "auto")
eng.precond = os.environ.get("VPRECOND", "twolevel")
only = os.environ.get("VGATES", "0,1,2,3")
seed = int(os.environ.get("VSEED", "2024")); nb = int(os.environ.get("VNETS", "3"))
rtol = float(os.environ.get("VRTOL", "1e-15"))
b = eng.generate([100000] * nb, 100.0, seed=seed)
res = eng.measure_fullnet(b, onoffmap=1, want_fields=True, rtol=rtol)
f = res["fields"]
sticks = eng.sticks_to_host(b)
gates = [("0", None, 0.0), ("back", "back", 10.0), ("total", "total", 10.0), ("partial", "partial", 10.0)]
print("iters", res["iters"].tolist())
for k in range(nb):
    s = sticks[k]
    jn = O.find_junctions(s)
    lab = O.label_clusters(len(s["xc"]), jn["stick1"], jn["stick2"])["label"]
    for q, (tag, gate, vgv) in enumerate(gates):
        if str(q) not in only.split(","): continue
        vg = O.gate_voltages(jn["x"], jn["y"], gate, vgv)
        cond = O.conductance(jn["kind2"], vg, 1, "linexp")
        r = O.solve_network(len(s["xc"]), jn["stick1"], jn["stick2"], cond, lab, refine=3)
        V = f["V"][q][b.h_soff[k]:b.h_soff[k + 1]].cpu().numpy()
        m = ~np.isnan(r["V"])
        err = np.abs(V[m] - r["V"][m]) / 0.1
        i = np.flatnonzero(m)[np.argmax(err)]
        deg = ((jn["stick1"] == i) | (jn["stick2"] == i))
        print("net %d gate %-7s max|dV|/Vds %.2e at stick %d (V gpu %.15f oracle %.15f) g of its junctions %s | vmax gpu %.3e over Vds; dI/I %.1e"
              % (k, tag, err.max(), i, V[i], r["V"][i], np.sort(cond[deg])[-4:].tolist(), V[m].max() - 0.1,
                 abs(res["isrc_all"][q, k, 2] - r["I_source"]) / r["I_source"]), flush=True)
