#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
: > gpurun_out/sweep.log
for cfg in "8 zero" "8 linear" "14 linear" "7 linear"; do
  set -- $cfg
  echo "== batch $1 guess $2" >> gpurun_out/sweep.log
  timeout 600 python bench.py --steps 2 --warmup 1 --batch $1 --guess $2 --no-cpu-baseline 2>> gpurun_out/sweep.err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']
    print('value %.2f e2e %.2f ms/step %.0f | pcg %.0f ms, %.0f iters, %.3f us/sys-iter | iters/solve %s'%(d['value'],d['e2e']['value'],d['ms_per_step'],r['pcg_seconds']*1e3,r['pcg_iterations'],r['pcg_seconds']*1e6/r['pcg_iterations'],[round(x) for x in d['pcg_iters_mean_per_solve']]))
" >> gpurun_out/sweep.log
done
cat gpurun_out/sweep.log; tail -3 gpurun_out/sweep.err
