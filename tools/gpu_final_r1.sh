#!/bin/bash
# final check of the round: full GPU suite, smoke, default bench, launch list of the default bench command
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 --timeout-method=thread -p no:cacheprovider --durations=6 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit: $?" >> gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_direct.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?" >> gpurun_out/ncu_launches.log
tail -12 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cut -c1-250 gpurun_out/bench.log; tail -2 gpurun_out/bench.err; tail -1 gpurun_out/ncu_launches.log; wc -l gpurun_out/launches_direct.csv
