#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_slab.py -q --timeout=600 --timeout-method=thread -p no:cacheprovider --durations=8 2>&1 | tail -80 > gpurun_out/pytest_slab.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_slab.log
tail -70 gpurun_out/pytest_slab.log
