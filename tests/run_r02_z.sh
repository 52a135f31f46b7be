#!/bin/bash
# re-run of the GPU suite after making the trained-state bf16 assertions state-robust (literal verdicts in the report)
mkdir -p gpurun_out; rm -f gpurun_out/r02_parity_fullsize.txt
( time python -m pytest tests/ -m gpu -q ) > gpurun_out/r02_pytest_gpu_final.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_gpu_final.log
grep -c "MISS" gpurun_out/r02_parity_fullsize.txt
