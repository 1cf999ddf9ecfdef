#!/bin/bash
# ncu evidence for the bench command: (1) launch list, (2) full capture of the top kernel's bench launch
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/ncu_plain.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?" >> gpurun_out/ncu_launches.log
timeout 600 $CMD > gpurun_out/ncu_plain2.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_pcg_cluster -s 1 -c 1 -o gpurun_out/prof_cluster_bench -f $CMD > gpurun_out/ncu_cluster.log 2>&1
echo "ncu exit $?" >> gpurun_out/ncu_cluster.log
tail -2 gpurun_out/ncu_launches.log; tail -2 gpurun_out/ncu_cluster.log; wc -l gpurun_out/launches.csv
