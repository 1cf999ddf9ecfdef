#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
: > gpurun_out/batch_sweep.log
for B in 14 21 28 42 56; do
  echo "== batch $B" >> gpurun_out/batch_sweep.log
  timeout 300 python bench.py --steps 3 --warmup 3 --batch $B --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.2f e2e %.2f ms/step %.1f | pcg_kernel %.1f ms | iters %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['stages_ms_rank0']['pcg_kernel'], d['roofline']['pcg_iterations']))" >> gpurun_out/batch_sweep.log
done
cat gpurun_out/batch_sweep.log
