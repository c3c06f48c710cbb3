#!/bin/bash
# full GPU suite on the candidate build, then C2 / C5 / whole-step A/B against _lib/base
mkdir -p gpurun_out
L=$PWD/lunaris_orion_b200/_lib
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_tests14_full.log 2>&1
tail -n 4 gpurun_out/r2_tests14_full.log
for v in base cur; do
  lib=$L/$v/liblunaris_b200.so; [ $v = cur ] && lib=$L/liblunaris_b200.so
  LUNARIS_B200_LIB=$lib PROF_TOP=12 timeout 300 python tools/profile_step.py 16 256 128 256 > gpurun_out/r2_prof14_c2_$v.log 2>&1
  echo "== $v C2"; grep -E "attn_fold|total kernel|affine|blk_bwd" gpurun_out/r2_prof14_c2_$v.log
done
for rep in 1 2; do
for v in base cur; do
  lib=$L/$v/liblunaris_b200.so; [ $v = cur ] && lib=$L/liblunaris_b200.so
  LUNARIS_B200_LIB=$lib timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench14_${v}_$rep.json 2> gpurun_out/r2_bench14_${v}_$rep.err
done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench14_*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, d['value'], d['e2e']['value'], d['value_repeat_after_e2e']['value'], d['c5']['value'], d['c2']['value'])
PY
