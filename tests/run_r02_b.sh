#!/bin/bash
# round 2, second GPU call: full GPU suite + bench (all workloads) with the fixed reference import
mkdir -p gpurun_out; rm -f gpurun_out/r02_parity_fullsize.txt
( time python -m pytest tests/ -m gpu -q -x ) > gpurun_out/r02_pytest_gpu_b.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_gpu_b.log
( time python bench.py --steps 50 --warmup 10 ) > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err
echo "bench rc=$?"
python bench.py --workload baseline --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_baseline_b.json 2> gpurun_out/r02_bench_baseline_b.err
echo "baseline rc=$?"
python bench.py --workload upperbound --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_upper_b.json 2> gpurun_out/r02_bench_upper_b.err
echo "upper rc=$?"
python bench.py --bn eval --steps 30 --warmup 5 --no-cpu-baseline --no-same-box > gpurun_out/r02_bench_evalbn_b.json 2> gpurun_out/r02_bench_evalbn_b.err
echo "evalbn rc=$?"
