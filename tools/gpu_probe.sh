#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
PROBE_METHODS=cluster PROBE_PRECOND=${PROBE_PRECOND:-twolevel} timeout 900 python tools/pcg_probe.py > gpurun_out/probe.log 2>&1; echo "exit $?" >> gpurun_out/probe.log
cat gpurun_out/probe.log
