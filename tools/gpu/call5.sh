#!/bin/bash
# GPU call: full GPU suite (complete log), bench, detailed step profile
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_tests5_full.log 2>&1
tail -n 40 gpurun_out/r2_tests5_full.log > gpurun_out/r2_tests5.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err
PROF_TOP=70 timeout 300 python tools/profile_step.py 64 > gpurun_out/r2_prof5_c3.log 2>&1
PROF_TOP=40 timeout 300 python tools/profile_step.py 16 256 128 256 > gpurun_out/r2_prof5_c2.log 2>&1
timeout 300 python tools/profile_sample.py > gpurun_out/r2_prof5_c5.log 2>&1
tail -n 12 gpurun_out/r2_tests5.log; tail -c 800 gpurun_out/r2_bench5.json
