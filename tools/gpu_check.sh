#!/bin/bash
# GPU run: parity tests (no -x, per-test timeout) + smoke + short bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout=900 --timeout-method=thread -p no:cacheprovider --durations=8 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit: $?" >> gpurun_out/smoke.log
timeout 1200 python bench.py --steps ${BENCH_STEPS:-2} --warmup ${BENCH_WARMUP:-1} --batch ${BENCH_BATCH:-8} > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit: $?" >> gpurun_out/bench.err
tail -60 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; cat gpurun_out/bench.log; tail -20 gpurun_out/bench.err
