#!/bin/bash
# round 2, session 2: spmv / currents / truth tests, bench with one and two engines, extra legs, 10M-stick network
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spmv.py tests/test_gpu_api.py tests/test_gpu_parity.py tests/test_gpu_ensemble.py -m gpu -q --timeout=600 -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/pytest_b.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_e1.log 2> gpurun_out/bench_e1.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra --engines 2 > gpurun_out/bench_e2.log 2> gpurun_out/bench_e2.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_extra.log 2> gpurun_out/bench_extra.err
BIG_SIDE=1000 BIG_CHECK=0 timeout 400 python tools/big_network.py > gpurun_out/big1000.log 2>&1
cat gpurun_out/pytest_b.log; tail -3 gpurun_out/bench_e1.err gpurun_out/bench_e2.err gpurun_out/bench_extra.err; cat gpurun_out/big1000.log
python - <<'PY'
import json
for f in ("bench_e1","bench_e2","bench_extra"):
    try:
        d=json.loads(open('gpurun_out/%s.log'%f).read().strip().splitlines()[-1])
        print(f, d['value'], d['e2e']['value'], d['ms_per_step'], {k:round(v/d['steps'],2) for k,v in d['stages_ms_rank0'].items()})
        if d.get('spmv'): print(json.dumps(d['spmv'])); print(json.dumps(d['configs'], indent=1))
    except Exception as e: print(f, "ERR", e)
PY
