#!/bin/bash
# same-box A/B of a candidate library (_lib/next) against the in-tree build: parity tests on the candidate, then whole-step
# bench lines alternating between the two builds
mkdir -p gpurun_out
L=$PWD/lunaris_orion_b200/_lib
(LUNARIS_B200_LIB=$L/next/liblunaris_b200.so timeout 900 python -m pytest tests/test_teacher_gpu.py tests/test_vae_gpu.py tests/test_c3_gpu.py tests/test_dropout_parity_gpu.py tests/test_fullsize_gpu.py tests/test_step_gpu.py -m gpu -q 2>&1 | tail -30) > gpurun_out/r2_tests8_next.log
for rep in 1 2; do
for v in cur next; do
  lib=$L/$v/liblunaris_b200.so; [ $v = cur ] && lib=$L/liblunaris_b200.so
  LUNARIS_B200_LIB=$lib timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench8_${v}_$rep.json 2> gpurun_out/r2_bench8_${v}_$rep.err
done
done
tail -n 4 gpurun_out/r2_tests8_next.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench8_*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, d['value'], d['e2e']['value'], d['value_repeat_after_e2e']['value'], d['c5']['value'], d['c2']['value'])
PY
