#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 800 python -m pytest tests/test_gpu_slab.py tests/test_gpu_parity.py -q --timeout=600 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -8
: > gpurun_out/rtol.log
for R in 1e-15 1e-14 1e-13 1e-12; do
  echo "== rtol $R" >> gpurun_out/rtol.log
  VRTOL=$R VNETS=2 timeout 300 python tools/v_error_probe.py 2>&1 | cut -c1-95,230- >> gpurun_out/rtol.log
done
cat gpurun_out/rtol.log
