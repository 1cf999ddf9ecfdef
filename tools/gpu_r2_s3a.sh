#!/bin/bash
# session 3: GPU suite + bench (no extras) after a kernel change
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout=600 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/pytest_s3.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_s3.log
timeout 600 python bench.py --no-cpu-baseline --no-extra --steps ${BENCH_STEPS:-6} --warmup 3 > gpurun_out/bench_s3.log 2> gpurun_out/bench_s3.err; echo "bench exit: $?" >> gpurun_out/bench_s3.err
tail -6 gpurun_out/pytest_s3.log; tail -3 gpurun_out/bench_s3.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_s3.log').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['steps'])
print({k: round(v/d['steps'],2) for k,v in d['stages_ms_rank0'].items()})
PY
