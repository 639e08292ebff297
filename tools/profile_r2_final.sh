#!/bin/bash
# Final round-2 captures of the kernels changed after tools/profile_r2.sh ran (one GPU; raw-metric CSV exports only).
set -u
OUT=gpurun_out
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base mangled"
cap () {   # name, mangled-name regex, launch count, command...
  local name=$1 regex=$2 count=$3; shift 3
  $NCU -k regex:"$regex" -c "$count" -f -o $OUT/$name "$@" > $OUT/$name.log 2>&1
  ncu -i $OUT/$name.ncu-rep --page raw --csv > $OUT/$name.csv 2>> $OUT/$name.log
  rm -f $OUT/$name.ncu-rep
}
cap r2f_fwd "spade_const_kernelILi3ELb0|spade_pixel_kernelILi3" 18 python bench.py --steps 1 --warmup 1 --no-graph --no-cpu --no-parity --no-train
cap r2f_gbwd "spade_const_kernelILi3ELb1|spade_wgrad_kernel|bilinear_adjoint|spade_combine" 24 \
    python bench.py --workload C3 --train-batch 8 --train-split 1 --steps 1 --warmup 1
cap r2f_dbwd "conv_wgrad_kernelILi3|conv3x3_wgrad_halo|conv_small" 24 python bench.py --workload C3 --train-batch 8 --train-split 1 --steps 1 --warmup 1
cap r2f_ops "upfirdn2d_sep_kernel" 8 python tools/microbench.py --iters 1
ls -la $OUT | grep r2f_
