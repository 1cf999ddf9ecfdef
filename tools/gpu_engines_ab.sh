#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for cfg in "$@"; do
  e=${cfg%%x*}; b=${cfg##*x}
  timeout 400 python bench.py --no-cpu-baseline --no-extra --steps 12 --warmup 4 --engines $e --batch $b > gpurun_out/ab.log 2> gpurun_out/ab.err || { echo "$cfg FAILED"; tail -5 gpurun_out/ab.err; continue; }
  python - "$cfg" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/ab.log').read().strip().splitlines()[-1])
print(sys.argv[1], round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],2))
PY
done
