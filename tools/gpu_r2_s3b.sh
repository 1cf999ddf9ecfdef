#!/bin/bash
# session 3: GPU suite + bench (no extras) + launch list with DRAM bytes
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
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches_s3.csv $CMD > gpurun_out/ncu_launches_s3.log 2>&1
echo "launch list exit $?"
