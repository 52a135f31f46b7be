#!/bin/bash
# usage: run_ncu_kernel.sh <kernel regex> <skip> <count> <out name>
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c $3 -o gpurun_out/$4 -f $CMD > gpurun_out/ncu_$4.log 2>&1
echo "capture exit $?"
