#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1500 --csv --log-file gpurun_out/launches2.csv $CMD > gpurun_out/ncu_launches2.log 2>&1
echo "launch list exit $?" >> gpurun_out/ncu_launches2.log
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:k_tiles|k_place|k_prep|k_bin" -c 4 -o gpurun_out/r02_tiles -f $CMD > gpurun_out/ncu_tiles.log 2>&1
echo "ncu tiles exit $?" >> gpurun_out/ncu_tiles.log
tail -n 2 gpurun_out/ncu_launches2.log gpurun_out/ncu_tiles.log
