#!/bin/bash
# ncu capture of the blocked GEMM's data-gradient variant (spade_const_kernel<3, true>) inside a training iteration.
set -u
OUT=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
$NCU --kernel-name-base mangled -k regex:"spade_const_kernelILi3ELb1" -c 4 -f -o $OUT/r2_prof_gbwd \
    python bench.py --workload C3 --train-batch 4 --train-split 1 --steps 1 --warmup 1 > $OUT/r2_prof_gbwd.log 2>&1
ncu -i $OUT/r2_prof_gbwd.ncu-rep --page raw --csv > $OUT/r2_prof_gbwd.csv 2>> $OUT/r2_prof_gbwd.log
ncu -i $OUT/r2_prof_gbwd.ncu-rep --page source --csv --print-source sass > $OUT/r2_prof_gbwd_sass.csv 2>> $OUT/r2_prof_gbwd.log
ncu -i $OUT/r2_prof_gbwd.ncu-rep --page details > $OUT/r2_prof_gbwd_details.txt 2>> $OUT/r2_prof_gbwd.log
rm -f $OUT/r2_prof_gbwd.ncu-rep
ls -la $OUT/ | grep gbwd
