#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${NGPU:-2}
BIG_SIDE=${BIG_SIDE:-300} timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/big_network_dist.py > gpurun_out/big_dist.log 2> gpurun_out/big_dist.err; echo "exit $?" >> gpurun_out/big_dist.err
cat gpurun_out/big_dist.log; grep -v "^W\|^\*\|OMP_NUM" gpurun_out/big_dist.err | tail -5
