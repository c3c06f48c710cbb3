#!/bin/bash
# GPU call: the driver's round-end sequence on one box: GPU suite, smoke(), default bench, reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_tests6_full.log 2>&1
tail -n 15 gpurun_out/r2_tests6_full.log > gpurun_out/r2_tests6.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke6.log 2>&1
timeout 900 python bench.py > gpurun_out/r2_bench6_default.json 2> gpurun_out/r2_bench6_default.err
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_ref6.json 2> gpurun_out/r2_ref6.err
tail -n 4 gpurun_out/r2_tests6.log; tail -n 2 gpurun_out/r2_smoke6.log; tail -c 700 gpurun_out/r2_bench6.json; tail -c 500 gpurun_out/r2_ref6.json
