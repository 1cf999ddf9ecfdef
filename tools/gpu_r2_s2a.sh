#!/bin/bash
# round 2, session 2: fixed tests, per-round timing of the symbolic phase, configs 2/4/5 with the current default solver,
# ncu: k_tiles / k_uf_union full capture, DRAM bytes of every direct-stage launch of one batch
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_api.py::test_load_reference_dump tests/test_gpu_direct.py::test_full_size_truth tests/test_gpu_spmv.py tests/test_gpu_slab.py -m gpu -q -s --timeout=600 -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/pytest_fix.log
CNT_SM_TIMING=1 timeout 300 python bench.py --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/bench_timing.log 2> gpurun_out/bench_timing.err
BIG_SIDE=300 BIG_CHECK=0 timeout 300 python tools/big_network.py > gpurun_out/big300.log 2>&1
BIG_SIDE=1000 BIG_CHECK=0 timeout 400 python tools/big_network.py > gpurun_out/big1000.log 2>&1
SWEEP_CHECK=0 timeout 300 python tools/gatesweep_bench.py > gpurun_out/gatesweep.log 2>&1
timeout 300 python tools/ensemble20_bench.py > gpurun_out/ens20.log 2>&1
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_tiles|k_uf_union" -c 2 -o gpurun_out/r02_tiles_uf -f $CMD > gpurun_out/ncu_tiles.log 2>&1
echo "ncu tiles exit $?" >> gpurun_out/ncu_tiles.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:k_sm_|k_mark|k_select|k_incident|k_pivots|k_items|k_init|k_hoff|k_final" -c 700 --csv --log-file gpurun_out/direct_stage_dram.csv $CMD > gpurun_out/ncu_direct.log 2>&1
echo "ncu direct exit $?" >> gpurun_out/ncu_direct.log
cat gpurun_out/pytest_fix.log; tail -3 gpurun_out/bench_timing.err; cat gpurun_out/big300.log gpurun_out/big1000.log gpurun_out/gatesweep.log gpurun_out/ens20.log | tail -40; tail -2 gpurun_out/ncu_tiles.log gpurun_out/ncu_direct.log
