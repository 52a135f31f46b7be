#!/bin/bash
# round 2, first GPU call: new full-size parity tests + full-tile operator cases + bench (all workloads, short)
mkdir -p gpurun_out; rm -f gpurun_out/r02_parity_fullsize.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_smi.txt
nproc >> gpurun_out/r02_smi.txt
( time python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "full_tile" ) > gpurun_out/r02_pytest_fulltile.log 2>&1
echo "fulltile rc=$?"
( time python -m pytest tests/test_gpu_fullsize.py -m gpu -q -s ) > gpurun_out/r02_pytest_fullsize.log 2>&1
echo "fullsize rc=$?"
tail -5 gpurun_out/r02_pytest_fullsize.log
( time python bench.py --steps 50 --warmup 10 ) > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err
echo "bench rc=$?"
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r02_bench_ref_a.json 2> gpurun_out/r02_bench_ref_a.err
echo "ref rc=$?"
python bench.py --workload baseline --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_baseline_a.json 2> gpurun_out/r02_bench_baseline_a.err
echo "baseline rc=$?"
python bench.py --workload upperbound --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_upper_a.json 2> gpurun_out/r02_bench_upper_a.err
echo "upper rc=$?"
