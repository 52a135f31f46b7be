mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-profile-pass"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"conv3x3|scribble_loss" -c 4000 --csv --log-file gpurun_out/traffic.csv $CMD > gpurun_out/ncu_traffic.log 2>&1
echo "traffic exit $?"
ncu --set full --clock-control none --import-source on -k regex:"scribble_loss_(fwd|bwd)" -s 2 -c 2 -o gpurun_out/prof_loss -f $CMD > gpurun_out/ncu_prof_loss.log 2>&1
ncu -i gpurun_out/prof_loss.ncu-rep --page raw --csv > gpurun_out/prof_loss_raw.csv 2>/dev/null
echo "loss capture exit $?"
python tests/timeline_profile.py 2>&1 | tail -6
