#!/bin/bash
# GPU call: new boundary tests, the stock-eager reference on the GPU (existing-kernel bar), launch list of two whole steps
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_boundary_gpu.py tests/test_heads_gpu.py -m gpu -q 2>&1 | tail -30) > gpurun_out/r2_tests7.log
timeout 900 python bench.py --impl reference --ref-device cuda --steps 2 --warmup 1 --eager-batch 16 > gpurun_out/r2_ref_eager_b16.json 2> gpurun_out/r2_ref_eager_b16.err
timeout 900 python bench.py --impl reference --ref-device cuda --steps 1 --warmup 1 --eager-batch 64 > gpurun_out/r2_ref_eager_b64.json 2> gpurun_out/r2_ref_eager_b64.err
python tools/prof_step_once.py 2 > gpurun_out/ncu_plain_step.log 2>&1 &&
ncu --clock-control none --metrics gpu__time_duration.sum -c 2600 --csv --log-file gpurun_out/r2_ncu_launches_2steps.csv python tools/prof_step_once.py 2 > gpurun_out/ncu_launches.log 2>&1
tail -n 5 gpurun_out/r2_tests7.log; cat gpurun_out/r2_ref_eager_b16.json gpurun_out/r2_ref_eager_b64.json | cut -c1-600; tail -n 3 gpurun_out/r2_ref_eager_b64.err; wc -l gpurun_out/r2_ncu_launches_2steps.csv
