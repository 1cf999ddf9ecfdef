#!/bin/bash
# round 2: the hand-written symbolic phase -- direct-solver tests, then a short bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_direct.py -m gpu -x -q --timeout=600 --timeout-method=thread -p no:cacheprovider --durations=8 2>&1 | tail -80 > gpurun_out/pytest_direct.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_direct.log
tail -30 gpurun_out/pytest_direct.log
timeout 600 python tools/dense_max_sweep.py 2>&1 | tee gpurun_out/dense_sweep.log
timeout 600 python bench.py --steps ${BENCH_STEPS:-5} --warmup ${BENCH_WARMUP:-3} --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit: $?" >> gpurun_out/bench.err
cat gpurun_out/bench.log; tail -20 gpurun_out/bench.err
