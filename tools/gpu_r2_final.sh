#!/bin/bash
# round 2 evidence run: GPU suite, smoke, default bench (+ reference arm), launch list, ncu captures
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 --timeout-method=thread -p no:cacheprovider --durations=5 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit: $?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err; echo "bench exit: $?" >> gpurun_out/bench_final.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2> gpurun_out/bench_reference.err
CMD="python bench.py --engines 1 --steps 1 --warmup 1 --no-cpu-baseline --no-extra"   # one stream: ncu serialises the launches anyway
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?" >> gpurun_out/ncu_launches.log
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_tiles" -c 1 -o gpurun_out/r02_tiles_final -f $CMD > gpurun_out/ncu_tiles.log 2>&1
echo "ncu tiles exit $?" >> gpurun_out/ncu_tiles.log
tail -6 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; tail -2 gpurun_out/bench_final.err; head -c 600 gpurun_out/bench_reference.log; echo; tail -n 1 gpurun_out/ncu_launches.log gpurun_out/ncu_tiles.log
