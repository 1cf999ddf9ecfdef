#!/bin/bash
# ncu evidence: (1) launch list of a short bench run, (2) full capture of the top kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --batch 4 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/ncu_plain.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?" >> gpurun_out/ncu_launches.log
export PROBE_NETS=2 PROBE_MAXIT=200 PROBE_METHODS=cluster PROBE_PRECOND=twolevel
timeout 300 python tools/pcg_probe.py > gpurun_out/probe_small.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pcg_cluster -c 1 -o gpurun_out/prof_cluster -f python tools/pcg_probe.py > gpurun_out/ncu_cluster.log 2>&1
echo "ncu exit $?" >> gpurun_out/ncu_cluster.log
tail -2 gpurun_out/ncu_launches.log; tail -2 gpurun_out/ncu_cluster.log; wc -l gpurun_out/launches.csv
