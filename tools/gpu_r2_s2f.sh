#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_generate.py tests/test_gpu_api.py tests/test_gpu_ensemble.py -m gpu -x -q --timeout=600 -p no:cacheprovider 2>&1 | tail -4 > gpurun_out/pytest_f.log
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_f.log 2> gpurun_out/bench_f.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_b224.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/ncu_launches.log 2>&1
cat gpurun_out/pytest_f.log; tail -3 gpurun_out/bench_f.err
python - <<'PY'
import json
for f in ("bench_f",):
    d=json.loads(open('gpurun_out/%s.log'%f).read().strip().splitlines()[-1])
    nb=d['config']['networks_per_gpu_per_step']/56.0
    print(f, d['value'], d['e2e']['value'], d['ms_per_step'], {k:round(v/d['steps']/nb,2) for k,v in d['stages_ms_rank0'].items()})
PY
