#!/bin/bash
# `--set full` captures of the fused scribble-loss kernels at the headline shape (12 pairs of 256^2, C = 5) and at the
# LVSC-scale shape (96 pairs of 224^2, C = 2); plain (un-profiled) runs first. Raw pages: gpurun_out/prof_loss*_raw.csv
mkdir -p gpurun_out
cap() {  # name, bench args
  local CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-profile-pass $2"
  $CMD > gpurun_out/plain_$1.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_$1.log; return 1; }
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"scribble_loss_(fwd|bwd)" -s 2 -c 2 \
      -o gpurun_out/$1 -f $CMD > gpurun_out/ncu_$1.log 2>&1
  echo "capture $1 exit $?"
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  wc -c gpurun_out/$1_raw.csv
}
cap prof_loss_r2 ""
cap prof_loss_lvsc "--batch 96 --size 224 --classes 2"
