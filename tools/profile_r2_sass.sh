#!/bin/bash
# One launch of each remaining training kernel under ncu with the SASS-level sampling exported (who stalls where).
set -u
OUT=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
one () {   # name, regex, skip
  local name=$1 regex=$2 skip=$3
  $NCU -k regex:"$regex" -s "$skip" -c 1 -f -o $OUT/r2_sass_$name \
      python bench.py --workload C3 --train-batch 8 --train-split 1 --steps 1 --warmup 1 > $OUT/r2_sass_$name.log 2>&1
  ncu -i $OUT/r2_sass_$name.ncu-rep --page raw --csv > $OUT/r2_sass_${name}_raw.csv 2>> $OUT/r2_sass_$name.log
  ncu -i $OUT/r2_sass_$name.ncu-rep --page source --csv --print-source sass > $OUT/r2_sass_${name}.csv 2>> $OUT/r2_sass_$name.log
  rm -f $OUT/r2_sass_$name.ncu-rep
}
one spade_wgrad "spade_wgrad_kernel" 3
one wgrad_halo "conv3x3_wgrad_halo_kernel" 8
one conv_halo "conv3x3_halo_kernel" 2
ls -la $OUT | grep r2_sass
