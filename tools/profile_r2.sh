#!/bin/bash
# Round-2 ncu captures (one GPU; numbers printed by programs under ncu are never bench values).  Writes gpurun_out/r2_prof_*.ncu-rep
set -u
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:"spade_const_kernel|spade_pixel_kernel|render_mlp_kernel|geo_kernel|geo_features" -c 24 -f -o gpurun_out/r2_prof_fwd \
    python bench.py --steps 1 --warmup 1 --no-graph --no-cpu --no-parity --no-train > gpurun_out/r2_prof_fwd.log 2>&1
$NCU -k regex:"conv3x3_halo_kernel|conv_kernel" -c 33 -f -o gpurun_out/r2_prof_dconv \
    python tools/dconv_layers.py > gpurun_out/r2_prof_dconv.log 2>&1
$NCU -k regex:"conv3x3_wgrad_halo_kernel|spade_const_kernel<3, true>|spade_wgrad_kernel|conv_wgrad_kernel" -c 14 -f -o gpurun_out/r2_prof_train \
    python bench.py --workload C3 --train-batch 4 --train-split 1 --steps 1 --warmup 1 > gpurun_out/r2_prof_train.log 2>&1
$NCU -k regex:"upfirdn2d_sep_kernel|bias_act_kernel" -c 6 -f -o gpurun_out/r2_prof_ops \
    python tools/microbench.py --iters 1 > gpurun_out/r2_prof_ops.log 2>&1
ls -la gpurun_out/*.ncu-rep
