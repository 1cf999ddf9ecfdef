#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ensemble.py tests/test_gpu_api.py -m gpu -q -x --timeout=600 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -5
timeout 600 python bench.py --no-cpu-baseline --no-extra > gpurun_out/bench_s3.log 2> gpurun_out/bench_s3.err; echo "bench exit: $?"; tail -3 gpurun_out/bench_s3.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_s3.log').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['steps'], d['config']['networks_per_gpu_per_step'], d['gpu_launches'])
print(d['stages_batches'], {k: round(v/d['stages_batches'],2) for k,v in d['stages_ms_rank0'].items()})
print({k: round(v/d['steps']/2,2) for k,v in d['stages_ms_two_streams_rank0'].items()})
r=d['roofline']; print(r['achieved'], r['frac'], r['traffic'], r['timed_region'], r['symbolic_ms_per_batch'], r['numeric_ms_per_batch'])
print(d['junction_tests_per_s'], d['junction_roofline']['frac'], d['allocator_cudaMalloc_calls'])
PY
