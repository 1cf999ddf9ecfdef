#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export PROBE_NETS=2 PROBE_MAXIT=200 PROBE_METHODS=cluster PROBE_PRECOND=twolevel
timeout 300 python tools/pcg_probe.py > gpurun_out/probe_small.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pcg_cluster -c 1 -o gpurun_out/prof_cluster -f python tools/pcg_probe.py > gpurun_out/ncu_cluster.log 2>&1
echo "ncu exit $?" >> gpurun_out/ncu_cluster.log
tail -3 gpurun_out/ncu_cluster.log
