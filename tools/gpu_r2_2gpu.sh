#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${NGPU:-2}
BIG_SIDE=1000 BIG_MODE=slab timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/big_network_dist.py > gpurun_out/big_slab_${N}gpu.log 2>&1
BIG_SIDE=1000 BIG_CHECK=0 timeout 300 python tools/big_network.py > gpurun_out/big1000_1gpu.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --no-extra > gpurun_out/bench_${N}gpu.log 2> gpurun_out/bench_${N}gpu.err
grep -v "^W\|^\[W\|warn" gpurun_out/big_slab_${N}gpu.log | tail -8; tail -3 gpurun_out/big1000_1gpu.log; tail -3 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_${N}gpu.log') if l.startswith('{')][-1])
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['n_gpus'])
pass
PY
