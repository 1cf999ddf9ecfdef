#!/bin/bash
# default bench under torchrun on N GPUs (the driver's launch line), no extras
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${NGPU:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps ${STEPS:-5} --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/bench_${N}gpu.log 2> gpurun_out/bench_${N}gpu.err
echo "exit $?"; tail -3 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_${N}gpu.log') if l.startswith('{')][-1])
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['n_gpus'], d['config']['networks_per_gpu_per_step'])
PY
