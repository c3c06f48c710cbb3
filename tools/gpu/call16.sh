#!/bin/bash
# convT halo kernel with staged (coalesced) output stores: parity, microbench A/B against _lib/base, C5 A/B
mkdir -p gpurun_out
L=$PWD/lunaris_orion_b200/_lib
(timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_vae_gpu.py tests/test_fullsize_gpu.py tests/test_properties_gpu.py -m gpu -q -x 2>&1 | tail -12) > gpurun_out/r2_tests16.log
tail -n 6 gpurun_out/r2_tests16.log
for v in base cur; do
  lib=$L/$v/liblunaris_b200.so; [ $v = cur ] && lib=$L/liblunaris_b200.so
  echo "== $v" | tee -a gpurun_out/r2_convt_ab16.log
  LUNARIS_B200_LIB=$lib timeout 300 python tools/bench_convt.py 2>&1 | grep halo | tee -a gpurun_out/r2_convt_ab16.log
  LUNARIS_B200_LIB=$lib timeout 300 python tools/bench_sample.py 2>&1 | head -1 | tee -a gpurun_out/r2_convt_ab16.log
done
