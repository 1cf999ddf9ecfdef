#!/bin/bash
# N-GPU run of the bench (torchrun, NCCL all-gather of the statistics rows)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${NGPU:-2}
nvidia-smi -L > gpurun_out/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${N}gpu.log 2> gpurun_out/bench_${N}gpu.err; echo "exit $?" >> gpurun_out/bench_${N}gpu.err
wc -l gpurun_out/gpus.txt; cat gpurun_out/bench_${N}gpu.log; tail -3 gpurun_out/bench_${N}gpu.err
