#!/bin/bash
# ncu --set full of one launch of every heavy kernel of the bench command (round 0 of the symbolic / numeric phases)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_init|k_select|k_pivots|k_items_a|k_items_b|k_items_scan|k_items_c|k_incident|k_mark" -c 9 -o gpurun_out/r02_sym_round0 -f $CMD > gpurun_out/ncu_a.log 2>&1
echo "ncu a exit $?" >> gpurun_out/ncu_a.log
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_sm_update|k_sm_pivot|k_sm_edges0|k_sm_coupling|k_sm_dense" -c 5 -o gpurun_out/r02_num_round0 -f $CMD > gpurun_out/ncu_b.log 2>&1
echo "ncu b exit $?" >> gpurun_out/ncu_b.log
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_stats_sticks|k_uf_union|k_rs_scatter|k_values|k_place|k_tiles|k_uf_flatten|k_currents" -c 9 -o gpurun_out/r02_front -f $CMD > gpurun_out/ncu_c.log 2>&1
echo "ncu c exit $?" >> gpurun_out/ncu_c.log
tail -n 2 gpurun_out/ncu_a.log gpurun_out/ncu_b.log gpurun_out/ncu_c.log; ls -la gpurun_out/*.ncu-rep
