#!/bin/bash
# specialised block-tail backward kernels + attn_fold issue spreading: parity, element microbench and in-situ times A/B
mkdir -p gpurun_out
L=$PWD/lunaris_orion_b200/_lib
(timeout 1200 python -m pytest tests/test_teacher_gpu.py tests/test_dropout_parity_gpu.py tests/test_c3_gpu.py tests/test_c2_gpu.py tests/test_step_gpu.py -m gpu -q -x 2>&1 | tail -6) > gpurun_out/r2_tests17.log
tail -n 3 gpurun_out/r2_tests17.log
for v in base cur; do
  lib=$L/$v/liblunaris_b200.so; [ $v = cur ] && lib=$L/liblunaris_b200.so
  echo "== $v"
  LUNARIS_B200_LIB=$lib timeout 300 python tools/bench_elem.py 2>&1 | grep -E "backward|epilogue"
  LUNARIS_B200_LIB=$lib PROF_TOP=12 timeout 300 python tools/profile_step.py 64 > gpurun_out/r2_prof17_c3_$v.log 2>&1
  grep -E "attn_fold|total kernel|blk_bwd|affine_tail|proj_expand" gpurun_out/r2_prof17_c3_$v.log
done
for rep in 1 2; do
for v in base cur; do
  lib=$L/$v/liblunaris_b200.so; [ $v = cur ] && lib=$L/liblunaris_b200.so
  LUNARIS_B200_LIB=$lib timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench17_${v}_$rep.json 2> gpurun_out/r2_bench17_${v}_$rep.err
done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench17_*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, d['value'], d['e2e']['value'], d['value_repeat_after_e2e']['value'], d['c5']['value'], d['c2']['value'])
PY
