#!/bin/bash
# C5 work: VAE / decoder parity on the candidate (in-tree) build, then the sampling timeline and throughput of the
# previous build (_lib/base) and the candidate on the same box
mkdir -p gpurun_out
L=$PWD/lunaris_orion_b200/_lib
(timeout 900 python -m pytest tests/test_vae_gpu.py tests/test_fullsize_gpu.py tests/test_properties_gpu.py tests/test_c1_gpu.py -m gpu -q -x 2>&1 | tail -30) > gpurun_out/r2_tests9.log
for v in base cur base cur; do
  lib=$L/$v/liblunaris_b200.so; [ $v = cur ] && lib=$L/liblunaris_b200.so
  echo "== $v" >> gpurun_out/r2_c5_ab9.log
  LUNARIS_B200_LIB=$lib timeout 300 python tools/bench_sample.py 2>&1 | head -1 >> gpurun_out/r2_c5_ab9.log
done
LUNARIS_B200_LIB=$L/base/liblunaris_b200.so timeout 300 python tools/profile_sample.py > gpurun_out/r2_prof9_c5_base.log 2>&1
timeout 300 python tools/profile_sample.py > gpurun_out/r2_prof9_c5_cur.log 2>&1
tail -n 5 gpurun_out/r2_tests9.log; cat gpurun_out/r2_c5_ab9.log; tail -n 20 gpurun_out/r2_prof9_c5_cur.log
