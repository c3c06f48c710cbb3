#!/bin/bash
# GPU call: full GPU suite (complete log), same-box conv A/B (previous in-tree build vs current), bench, profile
mkdir -p gpurun_out
L=$PWD/lunaris_orion_b200/_lib
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_tests4_full.log 2>&1
tail -n 40 gpurun_out/r2_tests4_full.log > gpurun_out/r2_tests4.log
for rep in 1 2; do
for v in base cur; do
  lib=$L/$v/liblunaris_b200.so; [ $v = cur ] && lib=$L/liblunaris_b200.so
  LUNARIS_B200_LIB=$lib timeout 300 python tools/bench_conv.py 64 > gpurun_out/r2_conv4_${v}_$rep.log 2>&1
done
done
LUN_CONV_STG2=0 timeout 300 python tools/bench_conv.py 64 > gpurun_out/r2_conv4_cur_stg1.log 2>&1
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err
timeout 300 python tools/profile_step.py 64 > gpurun_out/r2_prof4_c3.log 2>&1
tail -n 25 gpurun_out/r2_tests4.log; tail -c 800 gpurun_out/r2_bench4.json
