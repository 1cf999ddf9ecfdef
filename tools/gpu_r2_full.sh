#!/bin/bash
# round 2: whole GPU suite + smoke + bench + launch list of the bench command
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout=900 --timeout-method=thread -p no:cacheprovider --durations=8 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit: $?" >> gpurun_out/smoke.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit: $?" >> gpurun_out/bench.err
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_b56.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?" >> gpurun_out/ncu_launches.log
tail -40 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err; tail -2 gpurun_out/ncu_launches.log
