#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
run() {
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra 2> gpurun_out/ab.err > gpurun_out/ab.log || tail -3 gpurun_out/ab.err
  python - "$1" <<'PY'
import json, sys
d=json.loads(open("gpurun_out/ab.log").read().strip().splitlines()[-1])
nb=d["config"]["networks_per_gpu_per_step"]/56.0
print(sys.argv[1], round(d["value"],1), {k:round(v/d["steps"]/nb,2) for k,v in d["stages_ms_rank0"].items() if k in ("junction_count","junction_fill")}, "fallbacks", d["junction_fallback_batches"], flush=True)
PY
}
for cfg in "8 256 256 512" "10 256 256 512" "10 256 256 768" "12 384 384 1024" "12 256 256 768" "8 512 256 512" "6 256 256 512"; do
  set -- $cfg
  CNT_TILE_FLAGS="-DCNT_TILE_C=$1 -DCNT_TILE_CT=$2 -DCNT_TILE_HOME=$3 -DCNT_TILE_HIT=$4" python -m networksim_cntfet_b200.build > /dev/null 2>&1 || echo build failed
  run "TILE_C=$1 CT=$2 HOME=$3 HIT=$4"
done
python -m networksim_cntfet_b200.build > /dev/null 2>&1
