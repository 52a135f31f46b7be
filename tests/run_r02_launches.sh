#!/bin/bash
# round 2: per-launch device times (ncu --metrics gpu__time_duration.sum) for the train-BN and eval-BN steps
mkdir -p gpurun_out
export PP_CONV_TUNE_FILE=/tmp/pp_tune.txt
for bn in train eval; do
  CMD="python bench.py --bn $bn --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --no-profile-pass --no-same-box"
  $CMD > gpurun_out/plain_$bn.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_$bn.log; exit 1; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r02_launches_$bn.csv $CMD > gpurun_out/ncu_launches_$bn.log 2>&1
  echo "launch list $bn exit $?"
  python tests/agg_launches.py gpurun_out/r02_launches_$bn.csv 4 > gpurun_out/r02_launches_${bn}_agg.txt
done
