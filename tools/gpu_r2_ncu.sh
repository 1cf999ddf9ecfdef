#!/bin/bash
# round 2: ncu --set full captures of the heavy kernels of one bench step (first round of the symbolic phase,
# junction cell kernel, numeric update, dense tail)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/ncu_plain.log 2>&1 || exit 1
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:k_cells|k_init|k_incident|k_items_a|k_items_c" -c 6 -o gpurun_out/r02_sym_round0 -f $CMD > gpurun_out/ncu_sym.log 2>&1
echo "ncu sym exit $?" >> gpurun_out/ncu_sym.log
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:k_sm_pivot|k_sm_update|k_sm_edges0|k_sm_coupling" -c 4 -o gpurun_out/r02_num_round0 -f $CMD > gpurun_out/ncu_num.log 2>&1
echo "ncu num exit $?" >> gpurun_out/ncu_num.log
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:k_sm_dense|k_uf_union|k_prune_round|k_select|k_mark" -c 6 -o gpurun_out/r02_misc -f $CMD > gpurun_out/ncu_misc.log 2>&1
echo "ncu misc exit $?" >> gpurun_out/ncu_misc.log
tail -2 gpurun_out/ncu_sym.log gpurun_out/ncu_num.log gpurun_out/ncu_misc.log
ls -la gpurun_out/*.ncu-rep
