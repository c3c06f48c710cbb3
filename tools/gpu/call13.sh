#!/bin/bash
# attn_fold L2 prefetch + small-shape grid cap + C5 tail: parity on the candidate, in-situ kernel times and whole-step A/B
mkdir -p gpurun_out
L=$PWD/lunaris_orion_b200/_lib
(timeout 1200 python -m pytest tests/test_teacher_gpu.py tests/test_dropout_parity_gpu.py tests/test_c3_gpu.py tests/test_c2_gpu.py tests/test_vae_gpu.py tests/test_fullsize_gpu.py tests/test_properties_gpu.py -m gpu -q -x 2>&1 | tail -30) > gpurun_out/r2_tests13.log
tail -n 4 gpurun_out/r2_tests13.log
for v in base cur; do
  lib=$L/$v/liblunaris_b200.so; [ $v = cur ] && lib=$L/liblunaris_b200.so
  LUNARIS_B200_LIB=$lib PROF_TOP=14 timeout 300 python tools/profile_step.py 64 > gpurun_out/r2_prof13_c3_$v.log 2>&1
  LUNARIS_B200_LIB=$lib PROF_TOP=14 timeout 300 python tools/profile_step.py 16 256 128 256 > gpurun_out/r2_prof13_c2_$v.log 2>&1
  echo "== $v C3"; grep -E "attn_fold|total kernel|proj_expand|affine|blk_bwd" gpurun_out/r2_prof13_c3_$v.log
  echo "== $v C2"; grep -E "attn_fold|total kernel|proj_expand|affine|blk_bwd" gpurun_out/r2_prof13_c2_$v.log
done
for rep in 1 2; do
for v in base cur; do
  lib=$L/$v/liblunaris_b200.so; [ $v = cur ] && lib=$L/liblunaris_b200.so
  LUNARIS_B200_LIB=$lib timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench13_${v}_$rep.json 2> gpurun_out/r2_bench13_${v}_$rep.err
done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench13_*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, d['value'], d['e2e']['value'], d['value_repeat_after_e2e']['value'], d['c5']['value'], d['c2']['value'])
PY
