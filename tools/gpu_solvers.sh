#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_solvers.py -q --timeout=400 --timeout-method=thread -p no:cacheprovider --durations=6 2>&1 | tail -80 > gpurun_out/pytest_solvers.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_solvers.log
timeout 900 python bench.py --steps 2 --warmup 1 --batch 8 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit: $?" >> gpurun_out/bench.err
tail -60 gpurun_out/pytest_solvers.log; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
