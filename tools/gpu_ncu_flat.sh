#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export PROBE_NETS=8 PROBE_MAXIT=40 PROBE_METHODS=flat PROBE_PRECOND=twolevel CNT_NO_GRAPH=1
timeout 300 python tools/pcg_probe.py > gpurun_out/probe_flat.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 250 -c 300 --csv --log-file gpurun_out/launches_flat.csv python tools/pcg_probe.py > gpurun_out/ncu_flat.log 2>&1
echo "exit $?" >> gpurun_out/ncu_flat.log; tail -2 gpurun_out/ncu_flat.log; tail -3 gpurun_out/probe_flat.log
