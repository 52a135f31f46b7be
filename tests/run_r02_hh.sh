#!/bin/bash
# last GPU minutes of the round: the trained-state cells of the two pacing configs on the final code + final assertions
mkdir -p gpurun_out; rm -f gpurun_out/r02_parity_fullsize.txt
timeout 130 python -m pytest -q tests/test_gpu_fullsize.py -k "trained_state and (config2 or config3)" 2>&1 | tail -4
grep -c "MISS" gpurun_out/r02_parity_fullsize.txt
