#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 --timeout-method=thread -p no:cacheprovider --durations=5 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit: $?" >> gpurun_out/bench.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra --batch 56 > gpurun_out/bench_b56.log 2> gpurun_out/bench_b56.err
tail -12 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
for f in ("bench","bench_b56"):
    try:
        d=json.loads(open('gpurun_out/%s.log'%f).read().strip().splitlines()[-1])
        nb=d['config']['networks_per_gpu_per_step']/56.0
        print(f, d['value'], d['e2e']['value'], d['ms_per_step'], {k:round(v/d['steps']/nb,2) for k,v in d['stages_ms_rank0'].items()})
        if d.get('spmv'): print(json.dumps(d['spmv'])[:300]); print(json.dumps(d['configs'])[:3000])
    except Exception as e: print(f, "ERR", e)
PY
