#!/bin/bash
# round 2: DRAM bytes + duration per conv / loss / BatchNorm launch, and --set full captures of the dominant kernels
mkdir -p gpurun_out
export PP_CONV_TUNE_FILE=/tmp/pp_tune.txt
CMD="python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e --no-profile-pass --no-same-box --no-other-bn"
$CMD > gpurun_out/plain_traffic.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_traffic.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"conv3x3|scribble_loss|bn_" -c 6000 --csv --log-file gpurun_out/r02_ncu_traffic.csv $CMD > gpurun_out/ncu_traffic.log 2>&1
echo "traffic exit $?"
python tests/summarize_traffic.py gpurun_out/r02_ncu_traffic.csv gpurun_out/r02_conv_traffic.json profiles/r02_ncu_traffic.csv | tee gpurun_out/r02_conv_traffic.txt
ncu --set full --clock-control none --import-source on -k regex:"conv3x3_wgrad_tc_kernel" -s 20 -c 3 -f -o gpurun_out/r02_prof_wgrad $CMD > gpurun_out/ncu_prof_wgrad.log 2>&1
ncu -i gpurun_out/r02_prof_wgrad.ncu-rep --page raw --csv > gpurun_out/r02_prof_wgrad_raw.csv 2>/dev/null
echo "wgrad capture exit $?"
