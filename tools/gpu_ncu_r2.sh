#!/bin/bash
# ncu evidence, series-elimination build: bench launch list + full capture of the cluster kernel;
# sharded solve on one GPU: plain timing, launch list, full capture of its SpMV
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/ncu_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_series.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?" >> gpurun_out/ncu_launches.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pcg_cluster -s 1 -c 1 -o gpurun_out/prof_cluster_series -f $CMD > gpurun_out/ncu_cluster.log 2>&1
echo "ncu exit $?" >> gpurun_out/ncu_cluster.log
SLAB_SIDE=300 SLAB_WORLD=1 timeout 300 python tools/slab_probe.py > gpurun_out/slab_probe.log 2>&1
SLAB_SIDE=300 SLAB_WORLD=8 timeout 300 python tools/slab_probe.py >> gpurun_out/slab_probe.log 2>&1
SLAB_SIDE=1000 SLAB_WORLD=1 SLAB_REPS=1 timeout 300 python tools/slab_probe.py >> gpurun_out/slab_probe.log 2>&1
echo "probe exit $?" >> gpurun_out/slab_probe.log
CNT_SLAB_GRAPH=0 SLAB_SIDE=100 SLAB_REPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 1300 --csv --log-file gpurun_out/launches_slab.csv python tools/slab_probe.py > gpurun_out/ncu_slab_launches.log 2>&1
echo "slab launch list exit $?" >> gpurun_out/ncu_slab_launches.log
CNT_SLAB_GRAPH=0 SLAB_SIDE=300 SLAB_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_slab_spmv -s 100 -c 1 -o gpurun_out/prof_slab_spmv -f python tools/slab_probe.py > gpurun_out/ncu_slab.log 2>&1
echo "slab ncu exit $?" >> gpurun_out/ncu_slab.log
tail -2 gpurun_out/ncu_launches.log; tail -2 gpurun_out/ncu_cluster.log; cat gpurun_out/slab_probe.log; tail -2 gpurun_out/ncu_slab_launches.log; tail -2 gpurun_out/ncu_slab.log
