#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${NGPU:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "exit $?" >> gpurun_out/bench_${N}gpu.err
cat gpurun_out/bench_${N}gpu.json | cut -c1-400; grep -v "^W\|^\*\|OMP_NUM" gpurun_out/bench_${N}gpu.err | tail -5
