#!/bin/bash
# ncu evidence of the round: launch list of one step, --set full of the dominant conv kernels, of the bandwidth-bound
# Teacher passes and of the decoder-only sampling kernels. Every command runs plainly first.
mkdir -p gpurun_out
NCU="ncu --clock-control none"
python tools/prof_step_once.py 2 > gpurun_out/ncu_plain_step.log 2>&1 &&
$NCU --metrics gpu__time_duration.sum -s 560 -c 700 --csv --log-file gpurun_out/r2_ncu_launches_step.csv python tools/prof_step_once.py 2 > gpurun_out/ncu_launches.log 2>&1
python tools/prof_conv.py 512 512 3 64 stats+wgrad > gpurun_out/ncu_plain_conv.log 2>&1 &&
$NCU --set full --import-source on -k regex:"conv_fprop_kernel|conv_wgrad_kernel" -s 4 -c 2 -o gpurun_out/r2_ncu_conv_big python tools/prof_conv.py 512 512 3 64 stats+wgrad > gpurun_out/ncu_conv.log 2>&1
$NCU --set full --import-source on -k regex:"affine_fwd|attn_fold|proj_expand|blk_bwd_reduce|blk_bwd_apply" -s 40 -c 6 -o gpurun_out/r2_ncu_elem python tools/prof_step_once.py 1 > gpurun_out/ncu_elem.log 2>&1
python tools/profile_sample.py > gpurun_out/ncu_plain_c5.log 2>&1 &&
$NCU --set full --import-source on -k regex:"gn_mish_final|convt_halo|gn_mish_fwd" -s 12 -c 6 -o gpurun_out/r2_ncu_c5 python tools/profile_sample.py > gpurun_out/ncu_c5.log 2>&1
ls -la gpurun_out/*.ncu-rep
