#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python - > gpurun_out/occ.log 2>&1 <<'PY'
import torch
from networksim_cntfet_b200 import _lib
torch.zeros(1, device="cuda")
lib = _lib.load()
print({c: lib.cnt_pcg_cluster_active(c) for c in range(1, 17)})
PY
cat gpurun_out/occ.log
for c in 15 14; do
echo "== cluster size $c" >> gpurun_out/occ.log
CNT_CLUSTER_SIZE=$c PROBE_METHODS=cluster PROBE_PRECOND=twolevel timeout 300 python tools/pcg_probe.py 2>&1 | grep -E "twolevel cluster|total|Error|error" >> gpurun_out/occ.log
done
cat gpurun_out/occ.log
