#!/bin/bash
# per-launch device times of the BatchNorm kernels with and without replicated backward accumulators
mkdir -p gpurun_out
export PP_CONV_TUNE_FILE=/tmp/pp_tune.txt
CMD="python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --no-profile-pass --no-same-box"
$CMD --bn train > gpurun_out/plain_t.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_t.log; exit 1; }
for tag in reps1:train repsdef:train repsdef:eval; do
  r=${tag%%:*}; bn=${tag##*:}
  if [ $r = reps1 ]; then export PP_BN_REPLICAS=1; else unset PP_BN_REPLICAS; fi
  ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r02_t_launches_${r}_$bn.csv $CMD --bn $bn > gpurun_out/ncu_t_${r}_$bn.log 2>&1
  echo "launch list $tag exit $?"
  python tests/agg_launches.py gpurun_out/r02_t_launches_${r}_$bn.csv 4 > gpurun_out/r02_t_launches_${r}_${bn}_agg.txt
  grep -E "ms/step|bn_|maxpool|upsample" gpurun_out/r02_t_launches_${r}_${bn}_agg.txt
done
