#!/bin/bash
# GPU call: full GPU suite, bench (ours + reference arm), step profiles
mkdir -p gpurun_out
ls oracle/_ref > gpurun_out/r2_ref_files.txt 2>&1
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -60) > gpurun_out/r2_tests2.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
timeout 500 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_ref2.json 2> gpurun_out/r2_ref2.err
timeout 300 python tools/profile_step.py 64 > gpurun_out/r2_prof_c3.log 2>&1
timeout 300 python tools/profile_step.py 16 256 128 256 > gpurun_out/r2_prof_c2.log 2>&1
tail -n 5 gpurun_out/r2_tests2.log; tail -c 1500 gpurun_out/r2_bench2.json; tail -c 600 gpurun_out/r2_ref2.json
