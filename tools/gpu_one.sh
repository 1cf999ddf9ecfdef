#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest ${PYTEST_ARGS:-tests/test_gpu_ensemble.py} -q --timeout=600 --timeout-method=thread -p no:cacheprovider --durations=4 2>&1 | tail -60 > gpurun_out/pytest_one.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_one.log
tail -50 gpurun_out/pytest_one.log
