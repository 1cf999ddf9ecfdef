#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_direct.py -x -q --timeout=300 --timeout-method=thread -p no:cacheprovider --durations=5 2>&1 | tail -40 > gpurun_out/pytest_direct.log
cat gpurun_out/pytest_direct.log
timeout 300 python bench.py --steps 2 --warmup 2 --solver direct --no-cpu-baseline > gpurun_out/bench_direct.json 2> gpurun_out/bench_direct.err; echo "bench exit $?"
tail -3 gpurun_out/bench_direct.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_direct.json').read().strip().splitlines()[-1])
    print('value %.2f e2e %.2f ms/step %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step']), d['stages_ms_rank0'])
except Exception as e:
    print("no bench json", e)
PY
