#!/bin/bash
# A/B of library variants built into networksim_cntfet_b200/_ab/ (CNT_B200_LIB): bench without extras for each
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for lib in default $(ls networksim_cntfet_b200/_ab/*.so 2>/dev/null); do
  if [ "$lib" = default ]; then unset CNT_B200_LIB; else export CNT_B200_LIB=$PWD/$lib; fi
  timeout 300 python bench.py --no-cpu-baseline --no-extra --steps ${BENCH_STEPS:-5} --warmup 3 > gpurun_out/ab.log 2> gpurun_out/ab.err || { echo "$lib FAILED"; tail -5 gpurun_out/ab.err; continue; }
  python - "$lib" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/ab.log').read().strip().splitlines()[-1])
st={k: round(v/d['steps'],2) for k,v in d['stages_ms_rank0'].items()}
print(sys.argv[1], round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],2), st)
PY
done
