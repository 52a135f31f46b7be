#!/bin/bash
# Runs every case of build/test_conv_tc in its own process (bounded by `timeout`) and logs to gpurun_out/.
mkdir -p gpurun_out
LOG=gpurun_out/test_conv_tc.log
: > $LOG
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv >> $LOG 2>&1
N=$(./build/test_conv_tc -1)
rc_all=0
for i in $(seq 0 $((N-1))); do
  echo "=== case $i" >> $LOG
  timeout 120 ./build/test_conv_tc $i >> $LOG 2>&1
  rc=$?
  echo "exit $rc" >> $LOG
  if [ $rc -ne 0 ]; then rc_all=1; fi
done
grep -E "PASS|FAIL|exit|error|timeout" $LOG | tail -60
exit $rc_all
