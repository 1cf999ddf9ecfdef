#!/usr/bin/env python3
"""Row-sharded solve of one network with all ranks on ONE GPU (slab.LocalComm): per-iteration time of
the per-rank kernels without any interconnect, and their algorithmic HBM rate.
  SLAB_SIDE (um, default 300), SLAB_WORLD (ranks, default 1), CNT_SLAB_GRAPH=0 for eager launches."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from networksim_cntfet_b200.engine import Engine
from networksim_cntfet_b200 import slab
side = float(os.environ.get("SLAB_SIDE", "300")); world = int(os.environ.get("SLAB_WORLD", "1"))
n = int(10 * side * side)
eng = Engine(profile=True)
for rep in range(int(os.environ.get("SLAB_REPS", "2"))):
    b = eng.generate([n], side, seed=5)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = eng.measure_fullnet(b, onoffmap=1, row_shard=slab.LocalComm(world))
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    st = eng.stage_times_ms()
    its = int(eng.last_slab_info["issued"])
    NS = 4
    # algorithmic bytes of one iteration (all configurations): SpMV val 8 + col 4 + gathered p 8 per entry;
    # vectors: spmv p,diag r + q w (24), update x rw, p, q, r rw, dinv (56), aggregate sums r + member index (12),
    # direction r, dinv, S, einv, row_agg, p rw (44)  -> 20 B/entry + 136 B/row
    bytes_it = NS * (20.0 * b.NNZ + 136.0 * b.R)
    print("rep%d %gx%g um n=%d world %d (one GPU): %.2f s; rows %d nnz %d; iters %s; issued %d; slab_pcg %.1f ms -> %.1f us/iteration, %.0f GB/s algorithmic; graph %s"
          % (rep, side, side, n, world, dt, b.R, b.NNZ, res["iters"][:, 0].tolist(), its, st.get("slab_pcg", 0.0),
             1e3 * st.get("slab_pcg", 0.0) / max(its, 1), bytes_it * its / (st.get("slab_pcg", 1.0) * 1e-3) / 1e9,
             eng.last_slab_info["graph"]), flush=True)
    print("   ion %.12e back %.12e; stages ms %s" % (res["ion"][0], res["ioff_back"][0], {k: round(v, 1) for k, v in st.items()}), flush=True)
