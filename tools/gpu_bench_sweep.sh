#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
: > gpurun_out/sweep.log
for cfg in "flat 8" "flat 16" "flat 32" "cluster 16" "cluster 32"; do
  set -- $cfg
  echo "== solver $1 batch $2" >> gpurun_out/sweep.log
  timeout 600 python bench.py --steps 2 --warmup 1 --batch $2 --solver $1 --no-cpu-baseline 2>> gpurun_out/sweep.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']
    print('value %.2f e2e %.2f ms/step %.0f | pcg %.0f ms, %.0f iters, %.3f us/sys-iter, %.0f GB/s alg | stages %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],r['pcg_seconds']*1e3,r['pcg_iterations'],r['pcg_seconds']*1e6/r['pcg_iterations'],r['achieved'],d['stages_ms_rank0']))
" >> gpurun_out/sweep.log
done
cat gpurun_out/sweep.log; tail -3 gpurun_out/sweep.err
