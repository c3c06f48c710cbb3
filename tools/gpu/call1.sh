#!/bin/bash
# GPU call: full GPU suite on the in-tree library, bench (ours + reference arm), step profiles; then the candidate
# library (_lib/next) : kernel parity tests + A/B microbenchmarks against the in-tree and the round-1 builds
mkdir -p gpurun_out
L=$PWD/lunaris_orion_b200/_lib
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt; nproc >> gpurun_out/r2_gpu.txt
(timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_c3_gpu.py 2>&1 | tail -30) > gpurun_out/r2_tests1.log
(timeout 600 python -m pytest tests/test_c3_gpu.py -m gpu -q 2>&1 | tail -40) > gpurun_out/r2_tests1_c3.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_ref1.json 2> gpurun_out/r2_ref1.err
timeout 300 python tools/profile_step.py 64 > gpurun_out/r2_prof_c3.log 2>&1
timeout 300 python tools/profile_step.py 16 256 128 256 > gpurun_out/r2_prof_c2.log 2>&1
if [ -f $L/next/liblunaris_b200.so ]; then
  (LUNARIS_B200_LIB=$L/next/liblunaris_b200.so timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_teacher_gpu.py tests/test_vae_gpu.py tests/test_properties_gpu.py tests/test_fullsize_gpu.py -m gpu -q 2>&1 | tail -30) > gpurun_out/r2_tests1_next.log
  LUNARIS_B200_LIB=$L/next/liblunaris_b200.so timeout 200 python tools/gpu/check_head_linear.py > gpurun_out/r2_head_linear.log 2>&1
  for v in r1 base next; do
    lib=$L/$v/liblunaris_b200.so; [ $v = base ] && lib=$L/liblunaris_b200.so
    LUNARIS_B200_LIB=$lib timeout 300 python tools/bench_conv.py 64 > gpurun_out/r2_conv1_$v.log 2>&1
  done
  LUNARIS_B200_LIB=$L/next/liblunaris_b200.so LUN_CONV_STG2=0 timeout 300 python tools/bench_conv.py 64 > gpurun_out/r2_conv1_next_stg1.log 2>&1
  LUNARIS_B200_LIB=$L/next/liblunaris_b200.so LUN_CONV_STG2=2 timeout 300 python tools/bench_conv.py 64 > gpurun_out/r2_conv1_next_stg2forced.log 2>&1
  LUNARIS_B200_LIB=$L/next/liblunaris_b200.so LUN_WGRAD_LOCKSTEP=0 timeout 300 python tools/bench_conv.py 64 > gpurun_out/r2_conv1_next_nolock.log 2>&1
  LUNARIS_B200_LIB=$L/next/liblunaris_b200.so timeout 400 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_bench1_next.json 2> gpurun_out/r2_bench1_next.err
fi
tail -3 gpurun_out/r2_tests1.log gpurun_out/r2_tests1_c3.log gpurun_out/r2_tests1_next.log; tail -c 1200 gpurun_out/r2_bench1.json; tail -c 600 gpurun_out/r2_bench1_next.json
