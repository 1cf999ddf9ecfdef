#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for opt in "" "--engines 1 --batch 224"; do
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline $opt > gpurun_out/cfg4.log 2> gpurun_out/cfg4.err || { echo FAILED; tail -5 gpurun_out/cfg4.err; }
  python - "$opt" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/cfg4.log').read().strip().splitlines()[-1])
print(sys.argv[1], round(d['value'],1), {k:{kk:vv for kk,vv in v.items() if kk in ('value','seconds','sweep_ms','device_build_ms','error')} for k,v in d['configs'].items()})
PY
done
