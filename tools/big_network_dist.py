#!/usr/bin/env python3
"""One large network spread over the ranks of a torchrun job: BIG_MODE=slab (default: junction search sharded by
spatial slab, junction rows exchanged over NCCL, direct solve replicated), BIG_MODE=rows (unknown rows of the PCG
sharded, halo exchange + all-reduces over NCCL) or BIG_MODE=gates (gate configurations sharded, <= 4 useful ranks)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import networksim_cntfet_b200 as pkg
from networksim_cntfet_b200 import measure_perc
from networksim_cntfet_b200.engine import Engine
eng = pkg.set_engine(Engine("cuda:%d" % local))
side = float(os.environ.get("BIG_SIDE", "300")); n = int(10 * side * side)
mode = os.environ.get("BIG_MODE", "slab")
eng.profile = True
for rep in range(int(os.environ.get("BIG_REPS", "3"))):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    df, res = measure_perc.measure_fullnet_distributed(n, side, seed=5, mode=mode)
    torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
    if dist.get_rank() == 0:
        print("rep%d world %d mode %s: %gx%g um n=%d in %.2f s; ion %.12e ioff %s iters %s solver %s" % (rep, dist.get_world_size(), mode, side, side, n, dt,
              res["ion"][0], ["%.12e" % v for v in df.ioff.tolist()], res["iters"][:, 0].tolist(), eng.last_solver), flush=True)
        print("   stages ms", {k: round(v, 1) for k, v in eng.stage_times_ms().items()}, flush=True)
    else:
        eng.stage_times_ms()
dist.destroy_process_group()
