#!/bin/bash
# 2-GPU run of the bench (torchrun, NCCL all-gather of the statistics rows) + the reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_2gpu.log 2> gpurun_out/bench_2gpu.err; echo "exit $?" >> gpurun_out/bench_2gpu.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "exit $?" >> gpurun_out/bench_ref.err
cat gpurun_out/gpus.txt; cat gpurun_out/bench_2gpu.log; tail -5 gpurun_out/bench_2gpu.err; cat gpurun_out/bench_ref.log; tail -3 gpurun_out/bench_ref.err
