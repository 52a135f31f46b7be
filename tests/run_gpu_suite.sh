#!/bin/bash
# Full GPU validation pass; everything is logged under gpurun_out/ (merged back by gpurun).
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
echo "== pytest -m gpu" 
timeout 2400 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider -o faulthandler_timeout=600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
echo "== bench"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" | tee -a gpurun_out/bench.err
tail -3 gpurun_out/bench.log; tail -15 gpurun_out/bench.err
