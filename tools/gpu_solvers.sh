#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_solvers.py -q --timeout=240 --timeout-method=thread -p no:cacheprovider --durations=8 2>&1 | tail -100 > gpurun_out/pytest_solvers.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_solvers.log
tail -70 gpurun_out/pytest_solvers.log
