#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:k_tiles" -c 1 -o gpurun_out/r02_tiles2 -f $CMD > gpurun_out/ncu_tiles2.log 2>&1
echo "ncu tiles exit $?" >> gpurun_out/ncu_tiles2.log
tail -n 2 gpurun_out/ncu_tiles2.log
