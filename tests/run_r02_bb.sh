#!/bin/bash
# final launch list (train-BN headline step) + ncu --set full page of the final narrow weight-gradient kernel
mkdir -p gpurun_out
export PP_CONV_TUNE_FILE=/tmp/pp_tune.txt
CMD="python bench.py --bn train --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --no-profile-pass --no-same-box --no-other-bn"
$CMD > gpurun_out/plain_final.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_final.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_train_final.csv $CMD > gpurun_out/ncu_launches_train_final.log 2>&1
echo "launch list exit $?"
python tests/agg_launches.py gpurun_out/r02_launches_train_final.csv 4 > gpurun_out/r02_launches_train_final_agg.txt; head -30 gpurun_out/r02_launches_train_final_agg.txt
python tests/bench_wgrad_narrow.py > /dev/null 2>&1 || { echo "plain wgrad bench failed"; exit 1; }
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"conv3x3_wgrad_rowsn_tc_kernel<.int.32, .int.32>" -s 3 -c 1 \
  -o gpurun_out/r02_prof_wgrad_rowsn32 -f python tests/bench_wgrad_narrow.py > gpurun_out/ncu_r02_prof_wgrad_rowsn32.log 2>&1
echo "capture exit $?"
ncu -i gpurun_out/r02_prof_wgrad_rowsn32.ncu-rep --page raw --csv > gpurun_out/r02_ncu_prof_wgrad_rowsn32_final_raw.csv 2>/dev/null
wc -c gpurun_out/r02_ncu_prof_wgrad_rowsn32_final_raw.csv
