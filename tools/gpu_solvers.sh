#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_solvers.py -q --timeout=400 --timeout-method=thread -p no:cacheprovider --durations=4 2>&1 | tail -60 > gpurun_out/pytest_solvers.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_solvers.log
PROBE_METHODS=${PROBE_METHODS:-cluster} PROBE_PRECOND=${PROBE_PRECOND:-twolevel} timeout 600 python tools/pcg_probe.py > gpurun_out/probe.log 2>&1; echo "exit $?" >> gpurun_out/probe.log
tail -40 gpurun_out/pytest_solvers.log; cat gpurun_out/probe.log
