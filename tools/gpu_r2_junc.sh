#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_generate.py tests/test_gpu_api.py -m gpu -x -q --timeout=600 --timeout-method=thread -p no:cacheprovider --durations=5 2>&1 | tail -60 > gpurun_out/pytest_junc.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_junc.log
tail -40 gpurun_out/pytest_junc.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit: $?" >> gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'])
print({k:round(v/d['steps'],2) for k,v in d['stages_ms_rank0'].items()})
print(d['junction_tests_per_s'], d['junction_roofline'])
PY
tail -5 gpurun_out/bench.err
