#!/bin/bash
# the driver's round-end sequence on the final tree: GPU suite, smoke(), default bench, reference arm (short)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2c_tests_full.log 2>&1
tail -n 3 gpurun_out/r2c_tests_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c_smoke.log 2>&1; tail -c 200 gpurun_out/r2c_smoke.log; echo
timeout 900 python bench.py > gpurun_out/r2c_bench_default.json 2> gpurun_out/r2c_bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c_ref.json 2> gpurun_out/r2c_ref.err
python - <<'PY'
import json
for f in ['gpurun_out/r2c_bench_default.json']:
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, d['value'], d['e2e']['value'], d['value_repeat_after_e2e']['value'], d['c5']['value'], d['c2']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['gpu_launches'])
for l in open('gpurun_out/r2c_ref.json'):
    if l.startswith('{'):
        d=json.loads(l); print('ref', d['value'], d['unit'], d['impl'], d['cpu_baseline']['kind'], d['cpu_baseline']['cores'])
PY
