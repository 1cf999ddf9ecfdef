#!/bin/bash
# end of round: GPU suite, smoke, default bench (all legs)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 --timeout-method=thread -p no:cacheprovider --durations=5 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit: $?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err; echo "bench exit: $?" >> gpurun_out/bench_final.err
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; tail -2 gpurun_out/bench_final.err
