#!/bin/bash
# GPU run: parity tests (no -x, per-test timeout) + smoke
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider --durations=15 2>&1 | tail -120 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit: $?" >> gpurun_out/smoke.log
tail -60 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log
