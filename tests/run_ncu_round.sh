#!/bin/bash
# Round profile set (B200_PROFILING.md recipe): plain run first, then (1) per-launch durations of every kernel,
# (2) DRAM traffic of every conv / loss launch (-> tests/summarize_traffic.py -> profiles/r01_conv_traffic.json),
# (3) `--set full` captures of the dominant kernels (tests/run_ncu_kernels.sh) and of the loss / BatchNorm-backward kernels.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-profile-pass"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"conv3x3|scribble_loss" -c 4000 --csv --log-file gpurun_out/traffic.csv $CMD > gpurun_out/ncu_traffic.log 2>&1
echo "traffic exit $?"
cap() {  # name regex(base function name) skip count
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o gpurun_out/$1 -f $CMD > gpurun_out/ncu_$1.log 2>&1
  echo "capture $1 exit $?"
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
}
cap prof_loss "scribble_loss_(fwd|bwd)_kernel" 2 2
cap prof_bnbwd "bn_bwd_(apply|reduce)_kernel" 46 2
bash tests/run_ncu_kernels.sh
