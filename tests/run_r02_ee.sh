#!/bin/bash
# deterministic (fixed-point, integer atomics) aux-logit gradient: loss parity + the 2-GPU gradient-exchange check
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "scribble_loss or golden or data_parallel" 2>&1 | grep -E "^FAILED|passed|failed"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tests/run_dp_check.py > gpurun_out/r02_dp_check_2gpu_final.log 2>&1; echo "dp check exit $?"
grep -E "^step|^    |DP CHECK" gpurun_out/r02_dp_check_2gpu_final.log
timeout 300 $TR tests/run_dp_check.py 2>&1 | grep -E "^step|^    |DP CHECK"
