#!/bin/bash
# usage: run_ncu_list.sh [kernel-name regex] — per-launch durations (B200_PROFILING.md recipe) of one warm-up + two steps
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-profile-pass"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain.log; exit 1; }
FILTER=""
if [ -n "$1" ]; then FILTER="-k regex:$1"; fi
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 $FILTER --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
