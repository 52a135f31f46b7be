#!/bin/bash
# Profiling pass (B200_PROFILING.md recipe): plain run first, then the per-launch duration list and one
# `--set full` capture each of the tcgen05 conv forward/dgrad and wgrad kernels.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain.log; exit 1; }
tail -c 600 gpurun_out/plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc_kernel -s 60 -c 4 -o gpurun_out/prof_conv -f $CMD > gpurun_out/ncu_conv.log 2>&1
echo "conv capture exit $?"
ncu --set full --clock-control none --import-source on -k regex:conv3x3_wgrad_tc_kernel -s 30 -c 4 -o gpurun_out/prof_wgrad -f $CMD > gpurun_out/ncu_wgrad.log 2>&1
echo "wgrad capture exit $?"
ls -la gpurun_out/
