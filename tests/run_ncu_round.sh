#!/bin/bash
# Round profile set (B200_PROFILING.md recipe): plain run first, then (1) per-launch durations of every kernel,
# (2) DRAM traffic of every conv launch, (3) one `--set full` capture each of the dominant kernels.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-profile-pass"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"conv3x3|scribble_loss" -c 4000 --csv --log-file gpurun_out/traffic.csv $CMD > gpurun_out/ncu_traffic.log 2>&1
echo "traffic exit $?"
cap() {  # name regex skip count
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o gpurun_out/$1 -f $CMD > gpurun_out/ncu_$1.log 2>&1
  echo "capture $1 exit $?"
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
}
cap prof_conv256 "conv3x3_tc_kernel<256, 64>" 10 2
cap prof_wgrad256 "conv3x3_wgrad_tc_kernel<256, 2>" 6 2
cap prof_halo32 "conv3x3_halo_tc_kernel<32, 32>" 4 2
cap prof_loss "scribble_loss_(fwd|bwd)_kernel" 2 2
cap prof_bnbwd "bn_bwd_(apply|reduce)_kernel" 46 2
ls -la gpurun_out/ | head -40
