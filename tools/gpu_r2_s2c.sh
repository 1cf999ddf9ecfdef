#!/bin/bash
# batch-size sweep of the headline bench, config-2 ensemble through the direct solver, ncu capture of the SpMV
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for b in 112 224; do
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extra --batch $b > gpurun_out/bench_b$b.log 2> gpurun_out/bench_b$b.err
done
ENS_SOLVER=direct timeout 300 python tools/ensemble20_bench.py > gpurun_out/ens20_direct.log 2>&1
ENS_SOLVER=auto ENS_MAX_FILL=40 ENS_MAX_DEGREE=16 timeout 300 python tools/ensemble20_bench.py > gpurun_out/ens20_auto40.log 2>&1
cat > /tmp/spmv_only.py <<'PY'
import sys, json
sys.path.insert(0, ".")
import bench
from networksim_cntfet_b200.engine import Engine
print(json.dumps(bench.spmv_leg(Engine(), 600.0, reps=5)))
PY
timeout 300 python /tmp/spmv_only.py > gpurun_out/spmv_plain.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv_stream -s 3 -c 1 -o gpurun_out/r02_spmv -f python /tmp/spmv_only.py > gpurun_out/ncu_spmv.log 2>&1
echo "ncu spmv exit $?" >> gpurun_out/ncu_spmv.log
python - <<'PY'
import json
for f in ("bench_b112","bench_b224"):
    try:
        d=json.loads(open('gpurun_out/%s.log'%f).read().strip().splitlines()[-1])
        print(f, d['value'], d['e2e']['value'], d['ms_per_step'], {k:round(v/d['steps'],2) for k,v in d['stages_ms_rank0'].items()})
    except Exception as e: print(f, "ERR", e)
PY
tail -n 5 gpurun_out/bench_b224.err; cat gpurun_out/ens20_direct.log gpurun_out/ens20_auto40.log gpurun_out/spmv_plain.log; tail -n 2 gpurun_out/ncu_spmv.log
