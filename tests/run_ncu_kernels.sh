#!/bin/bash
# `--set full` captures of the dominant kernels (B200_PROFILING.md recipe; after a plain run exited 0 without ncu).
# Kernel templates are selected on the DEMANGLED name ("<(int)256, (int)64, (int)2>", matched with "." wildcards). Raw pages land in gpurun_out/<name>_raw.csv.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-profile-pass"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain.log; exit 1; }
cap() {  # name regex skip count
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$2" -s $3 -c $4 \
      -o gpurun_out/$1 -f $CMD > gpurun_out/ncu_$1.log 2>&1
  echo "capture $1 exit $?"
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  wc -c gpurun_out/$1_raw.csv
}
cap prof_conv256 "conv3x3_tc_kernel<.int.256, .int.64, .int.2>" 12 3
cap prof_conv128 "conv3x3_tc_kernel<.int.128, .int.64, .int.2>" 4 2
cap prof_wgrad256 "conv3x3_wgrad_tc_kernel<.int.256, .int.2>" 9 3
cap prof_halo32 "conv3x3_halo_tc_kernel<.int.32, .int.32>" 8 3
cap prof_wgrad_narrow "conv3x3_wgrad_narrow_tc_kernel<.int.32, .int.32>" 3 2
