#!/bin/bash
# in-situ kernel times (CUPTI) of one C3 step and the element microbench: previous build (_lib/base) vs candidate
mkdir -p gpurun_out
L=$PWD/lunaris_orion_b200/_lib
for v in base cur; do
  lib=$L/$v/liblunaris_b200.so; [ $v = cur ] && lib=$L/liblunaris_b200.so
  echo "== $v" >> gpurun_out/r2_elem_ab11.log
  LUNARIS_B200_LIB=$lib timeout 300 python tools/bench_elem.py 2>&1 | tail -8 >> gpurun_out/r2_elem_ab11.log
  LUNARIS_B200_LIB=$lib PROF_TOP=24 timeout 300 python tools/profile_step.py 64 > gpurun_out/r2_prof11_c3_$v.log 2>&1
done
cat gpurun_out/r2_elem_ab11.log
grep -E "affine|blk_bwd|ring|total kernel|proj_expand|attn_fold|channel_stats" gpurun_out/r2_prof11_c3_base.log
echo; grep -E "affine|blk_bwd|ring|total kernel|proj_expand|attn_fold|channel_stats" gpurun_out/r2_prof11_c3_cur.log
