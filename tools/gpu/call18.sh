#!/bin/bash
# evidence run of the second round-2 session: GPU suite, smoke(), bench lines, CUPTI step breakdowns, ncu launch list of
# one step and ncu --set full of the kernels changed in this session (every ncu command runs plainly first)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2b_tests_full.log 2>&1
tail -n 3 gpurun_out/r2b_tests_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1; tail -c 300 gpurun_out/r2b_smoke.log; echo
timeout 900 python bench.py > gpurun_out/r2b_bench_default.json 2> gpurun_out/r2b_bench_default.err
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench_c3_steps20.json 2> gpurun_out/r2b_bench_c3_steps20.err
PROF_TOP=70 timeout 300 python tools/profile_step.py 64 > gpurun_out/r2b_prof_c3.log 2>&1
PROF_TOP=40 timeout 300 python tools/profile_step.py 16 256 128 256 > gpurun_out/r2b_prof_c2.log 2>&1
timeout 300 python tools/profile_sample.py > gpurun_out/r2b_prof_c5.log 2>&1
NCU="ncu --clock-control none"
python tools/prof_step_once.py 2 > gpurun_out/ncu_plain_step.log 2>&1 &&
$NCU --metrics gpu__time_duration.sum -s 560 -c 700 --csv --log-file gpurun_out/r2b_ncu_launches_step.csv python tools/prof_step_once.py 2 > gpurun_out/ncu_launches.log 2>&1
$NCU --set full --import-source on -k regex:"affine_tail|attn_fold|proj_expand|blk_bwd_reduce_fast|blk_bwd_apply_fast" -s 40 -c 6 -f -o gpurun_out/r2b_ncu_elem python tools/prof_step_once.py 1 > gpurun_out/ncu_elem.log 2>&1
python tools/profile_sample.py > gpurun_out/ncu_plain_c5.log 2>&1 &&
$NCU --set full --import-source on -k regex:"convt_halo" -s 4 -c 2 -f -o gpurun_out/r2b_ncu_convt python tools/profile_sample.py > gpurun_out/ncu_c5.log 2>&1
ls -la gpurun_out/r2b_*.ncu-rep
python - <<'PY'
import json
for f in ['gpurun_out/r2b_bench_default.json','gpurun_out/r2b_bench_c3_steps20.json']:
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, d['value'], d['e2e']['value'], d['value_repeat_after_e2e']['value'], d['c5']['value'], d['c2']['value'], d['roofline']['frac'])
            print(json.dumps(d.get('hbm_passes'))[:1500])
PY
