#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python tools/gatesweep_bench.py > gpurun_out/gatesweep.log 2>&1; echo "exit $?" >> gpurun_out/gatesweep.log
timeout 600 python tools/ensemble20_bench.py > gpurun_out/ens20.log 2>&1; echo "exit $?" >> gpurun_out/ens20.log
cat gpurun_out/gatesweep.log gpurun_out/ens20.log
