#!/bin/bash
# ncu launch list of two whole C3 steps on the final build (the plain run first), summarised from the last step's first kernel
mkdir -p gpurun_out
python tools/prof_step_once.py 2 > gpurun_out/ncu_plain_step.log 2>&1 &&
ncu --clock-control none --metrics gpu__time_duration.sum -c 3400 --csv --log-file gpurun_out/r2b_ncu_launches_2steps.csv python tools/prof_step_once.py 2 > gpurun_out/ncu_launches.log 2>&1
python tools/summarize_launches.py gpurun_out/r2b_ncu_launches_2steps.csv --from-last conv3x3_c3_fwd | head -30
