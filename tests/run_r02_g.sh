#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x ) > gpurun_out/r02_pytest_g.log 2>&1
echo "parity tests rc=$?"; tail -2 gpurun_out/r02_pytest_g.log
PP_CONV_TUNE_DEBUG=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-same-box --no-e2e > gpurun_out/r02_bench_g.json 2> gpurun_out/r02_bench_g.err
echo "bench rc=$?"
timeout 300 python bench.py --bn eval --steps 30 --warmup 5 --no-cpu-baseline --no-same-box --no-e2e > gpurun_out/r02_bench_g_eval.json 2> gpurun_out/r02_bench_g_eval.err
echo "bench eval rc=$?"
PP_CONV_ROWS=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-same-box --no-e2e > gpurun_out/r02_bench_g_norows.json 2> gpurun_out/r02_bench_g_norows.err
echo "bench norows rc=$?"
