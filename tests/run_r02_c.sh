#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r02_parity_fullsize.txt
( time python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bn_eval or eval_bn or os16_eval or properties" ) > gpurun_out/r02_pytest_c1.log 2>&1
echo "eval tests rc=$?"; tail -3 gpurun_out/r02_pytest_c1.log
( time python -m pytest tests/ -m gpu -q ) > gpurun_out/r02_pytest_gpu_c.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_gpu_c.log
python bench.py --bn eval --steps 30 --warmup 5 --no-cpu-baseline --no-same-box > gpurun_out/r02_bench_evalbn_c.json 2> gpurun_out/r02_bench_evalbn_c.err
echo "evalbn rc=$?"
PP_NO_EVAL_FUSION=1 python bench.py --bn eval --steps 30 --warmup 5 --no-cpu-baseline --no-same-box > gpurun_out/r02_bench_evalbn_nofuse_c.json 2> gpurun_out/r02_bench_evalbn_nofuse_c.err
echo "evalbn nofuse rc=$?"
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-same-box > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err
echo "train rc=$?"
