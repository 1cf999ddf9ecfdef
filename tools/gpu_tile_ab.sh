#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
run() {
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra 2> gpurun_out/ab.err > gpurun_out/ab.log
  python - "$1" <<'PY'
import json, sys
d=json.loads(open("gpurun_out/ab.log").read().strip().splitlines()[-1])
nb=d["config"]["networks_per_gpu_per_step"]/56.0
print(sys.argv[1], round(d["value"],1), {k:round(v/d["steps"]/nb,2) for k,v in d["stages_ms_rank0"].items() if k in ("junction_count","junction_fill","direct_symbolic","direct_numeric","clusters")}, flush=True)
PY
}
run "default 512/2048/1024 (no line staging)"
for cfg in "256 1024 512" "256 2048 1024" "384 2048 768" "512 2048 512"; do
  set -- $cfg
  CNT_TILE_FLAGS="-DCNT_TILE_CT=$1 -DCNT_TILE_Q=$2 -DCNT_TILE_HIT=$3" python -m networksim_cntfet_b200.build > /dev/null 2>&1 || echo build failed
  run "CT=$1 Q=$2 HIT=$3"
done
python -m networksim_cntfet_b200.build > /dev/null 2>&1
