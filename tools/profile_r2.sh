#!/bin/bash
# Round-2 ncu captures (one GPU; numbers printed by programs under ncu are never bench values).
# Reports are large (--set full, tens of launches), gpurun brings back at most 64 MiB: every report is exported to CSV on the box
# (`--page raw`), only two single-launch reports (the dominant forward kernel, the haloed convolution) are kept as .ncu-rep.
set -u
OUT=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
cap () {   # name, kernel regex, launch count, command...
  local name=$1 regex=$2 count=$3; shift 3
  $NCU -k regex:"$regex" -c "$count" -f -o $OUT/$name "$@" > $OUT/$name.log 2>&1
  ncu -i $OUT/$name.ncu-rep --page raw --csv > $OUT/$name.csv 2>> $OUT/$name.log
  rm -f $OUT/$name.ncu-rep
}
cap r2_prof_fwd "spade_const_kernel|spade_pixel_kernel|render_mlp_kernel|geo_kernel|geo_features" 22 \
    python bench.py --steps 1 --warmup 1 --no-graph --no-cpu --no-parity --no-train
cap r2_prof_dconv "conv3x3_halo_kernel|conv_kernel" 33 python tools/dconv_layers.py
cap r2_prof_train "conv3x3_wgrad_halo_kernel|spade_const_kernel<3, true>|spade_wgrad_kernel|conv_wgrad_kernel|combine" 16 \
    python bench.py --workload C3 --train-batch 4 --train-split 1 --steps 1 --warmup 1
cap r2_prof_ops "upfirdn2d_sep_kernel|bias_act_kernel" 6 python tools/microbench.py --iters 1
# two single-launch reports kept for re-import (source view)
$NCU -k regex:"spade_const_kernel" -s 2 -c 1 -f -o $OUT/r2_rep_spade_const python bench.py --steps 1 --warmup 1 --no-graph --no-cpu --no-parity --no-train > /dev/null 2>&1
$NCU -k regex:"conv3x3_halo_kernel" -s 0 -c 1 -f -o $OUT/r2_rep_conv3x3_halo python tools/dconv_layers.py > /dev/null 2>&1
ls -la $OUT/
