#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
BIG_SIDE=300 timeout 900 python tools/big_network.py > gpurun_out/big.log 2>&1; echo "exit $?" >> gpurun_out/big.log
BIG_SIDE=1000 BIG_CHECK=0 timeout 900 python tools/big_network.py >> gpurun_out/big.log 2>&1; echo "exit $?" >> gpurun_out/big.log
cat gpurun_out/big.log | tail -30
