#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r02_parity_fullsize.txt
( time python -m pytest tests/ -m gpu -q ) > gpurun_out/r02_pytest_gpu_d.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_gpu_d.log
