#!/bin/bash
# one ncu --set full capture (with source) of the kernels matching $1 (regex), launch-skip $2, count $3, output name $4
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$1" -s ${2:-0} -c ${3:-1} -o gpurun_out/$4 -f $CMD > gpurun_out/ncu_$4.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_$4.log
