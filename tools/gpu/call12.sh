#!/bin/bash
# C5: parity of the candidate build, sampling throughput A/B against _lib/base, timeline, then ncu --set full of the tail
# and GroupNorm+Mish kernels (after the plain run exited 0)
mkdir -p gpurun_out
L=$PWD/lunaris_orion_b200/_lib
(timeout 900 python -m pytest tests/test_vae_gpu.py tests/test_fullsize_gpu.py tests/test_properties_gpu.py -m gpu -q -x 2>&1 | tail -5) > gpurun_out/r2_tests15.log
tail -n 3 gpurun_out/r2_tests15.log
for v in base cur base cur; do
  lib=$L/$v/liblunaris_b200.so; [ $v = cur ] && lib=$L/liblunaris_b200.so
  echo "== $v" >> gpurun_out/r2_c5_ab15.log
  LUNARIS_B200_LIB=$lib timeout 300 python tools/bench_sample.py 2>&1 | head -1 >> gpurun_out/r2_c5_ab15.log
done
cat gpurun_out/r2_c5_ab15.log
timeout 300 python tools/profile_sample.py > gpurun_out/r2_prof15_c5_cur.log 2>&1 && tail -n 18 gpurun_out/r2_prof15_c5_cur.log &&
ncu --clock-control none --set full --import-source on -k regex:"gn_mish_final|gn_mish_fwd" -s 8 -c 4 -o gpurun_out/r2b_ncu_c5_tail2 -f python tools/profile_sample.py > gpurun_out/ncu_c5_15.log 2>&1
ls -la gpurun_out/r2b_ncu_c5_tail2.ncu-rep
