#!/bin/bash
# GPU call: full GPU suite on the new kernels, conv A/B against the previous builds, bench, step profiles
mkdir -p gpurun_out
L=$PWD/lunaris_orion_b200/_lib
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -120) > gpurun_out/r2_tests3.log
for v in r1 base cur; do
  lib=$L/$v/liblunaris_b200.so; [ $v = cur ] && lib=$L/liblunaris_b200.so
  LUNARIS_B200_LIB=$lib timeout 300 python tools/bench_conv.py 64 > gpurun_out/r2_conv3_$v.log 2>&1
done
LUN_CONV_STG2=0 timeout 300 python tools/bench_conv.py 64 > gpurun_out/r2_conv3_cur_stg1.log 2>&1
LUN_WGRAD_LOCKSTEP=0 timeout 300 python tools/bench_conv.py 64 > gpurun_out/r2_conv3_cur_nolock.log 2>&1
timeout 200 python tools/gpu/check_head_linear.py > gpurun_out/r2_head_linear3.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
timeout 300 python tools/profile_step.py 64 > gpurun_out/r2_prof3_c3.log 2>&1
timeout 300 python tools/profile_step.py 16 256 128 256 > gpurun_out/r2_prof3_c2.log 2>&1
tail -n 8 gpurun_out/r2_tests3.log; tail -c 1500 gpurun_out/r2_bench3.json
