#!/bin/bash
# bench (default arguments) + ncu evidence of the same command: launch list + full capture of the cluster kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/ncu_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_b56.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?" >> gpurun_out/ncu_launches.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_pcg_cluster -s 1 -c 1 -o gpurun_out/prof_cluster_b56 -f $CMD > gpurun_out/ncu_cluster.log 2>&1
echo "ncu exit $?" >> gpurun_out/ncu_cluster.log
cat gpurun_out/bench.log; tail -3 gpurun_out/bench.err; tail -2 gpurun_out/ncu_launches.log; tail -2 gpurun_out/ncu_cluster.log
