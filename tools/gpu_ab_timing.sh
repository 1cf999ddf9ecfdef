#!/bin/bash
# A/B of library variants with the per-round kernel times of the symbolic phase (CNT_SM_TIMING=1)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for lib in default $(ls networksim_cntfet_b200/_ab/*.so 2>/dev/null); do
  if [ "$lib" = default ]; then unset CNT_B200_LIB; tag=default; else export CNT_B200_LIB=$PWD/$lib; tag=$(basename $lib .so); fi
  CNT_SM_TIMING=1 timeout 300 python bench.py --no-cpu-baseline --no-extra --steps 2 --warmup 2 > gpurun_out/abt_$tag.log 2> gpurun_out/abt_$tag.err || { echo "$lib FAILED"; tail -5 gpurun_out/abt_$tag.err; continue; }
  echo "== $tag"; grep "sm totals\|sm init" gpurun_out/abt_$tag.err | tail -2
done
