#!/bin/bash
# usage: run_dp_suite.sh N — multi-GPU validation: gradient-exchange check, then bench at 1 and N GPUs
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tests/run_dp_check.py > gpurun_out/dp_check_${N}gpu.log 2>&1; echo "dp check exit $?"
grep -E "step|DP CHECK|Error|error" gpurun_out/dp_check_${N}gpu.log | tail -8
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_1gpu.log 2> gpurun_out/bench_1gpu.err; echo "bench 1 exit $?"
tail -c 2500 gpurun_out/bench_1gpu.log
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu.log 2> gpurun_out/bench_${N}gpu.err; echo "bench $N exit $?"
tail -c 2500 gpurun_out/bench_${N}gpu.log; tail -5 gpurun_out/bench_${N}gpu.err
