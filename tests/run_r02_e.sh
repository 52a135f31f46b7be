#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "conv3x3_forward or full_tile or bn_eval_fused" ) > gpurun_out/r02_pytest_e1.log 2>&1
echo "conv op tests rc=$?"; tail -3 gpurun_out/r02_pytest_e1.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-same-box --no-e2e > gpurun_out/r02_bench_e.json 2> gpurun_out/r02_bench_e.err
echo "bench rows rc=$?"
PP_CONV_ROWS=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-same-box --no-e2e > gpurun_out/r02_bench_e_norows.json 2> gpurun_out/r02_bench_e_norows.err
echo "bench norows rc=$?"
timeout 300 python bench.py --bn eval --steps 30 --warmup 5 --no-cpu-baseline --no-same-box --no-e2e > gpurun_out/r02_bench_e_eval.json 2> gpurun_out/r02_bench_e_eval.err
echo "bench eval rc=$?"
