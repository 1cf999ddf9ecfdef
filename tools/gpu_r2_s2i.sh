#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_direct.py tests/test_gpu_ensemble.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_i.log 2> gpurun_out/bench_i.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_i.log').read().strip().splitlines()[-1])
nb=d['config']['networks_per_gpu_per_step']/56.0
print(d['value'], d['e2e']['value'], d['ms_per_step'], {k:round(v/d['steps']/nb,2) for k,v in d['stages_ms_rank0'].items()})
PY
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_stats_sticks|k_stats_second|k_uf_union|k_uf_flatten|k_values|k_place|k_currents|k_voltages|k_pattern|k_sort_rows|k_bin|k_prep|k_dense|k_gather|k_generate" -c 16 -o gpurun_out/r02_front2 -f $CMD > gpurun_out/ncu_d.log 2>&1
echo "ncu d exit $?" >> gpurun_out/ncu_d.log; tail -n 2 gpurun_out/ncu_d.log
