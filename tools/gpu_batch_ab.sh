#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for b in "$@"; do
  timeout 400 python bench.py --no-cpu-baseline --no-extra --steps 4 --warmup 3 --batch $b > gpurun_out/ab.log 2> gpurun_out/ab.err || { echo "$b FAILED"; tail -5 gpurun_out/ab.err; continue; }
  python - "$b" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/ab.log').read().strip().splitlines()[-1])
st={k: round(v/d['steps'],2) for k,v in d['stages_ms_rank0'].items()}
print(sys.argv[1], round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],2), d['allocator_cudaMalloc_calls']['reserved_GB'], st)
PY
done
