#!/bin/bash
mkdir -p gpurun_out
( PP_CONV_FORCE=rows timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "conv3x3_forward or full_tile or bn_eval_fused" ) > gpurun_out/r02_pytest_i.log 2>&1
echo "forced rows op tests rc=$?"; tail -2 gpurun_out/r02_pytest_i.log
( PP_CONV_FORCE=rows PP_CONV_ROWS_PAIR=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "conv3x3_forward or full_tile or bn_eval_fused" ) > gpurun_out/r02_pytest_i2.log 2>&1
echo "forced pair op tests rc=$?"; tail -2 gpurun_out/r02_pytest_i2.log
PP_ROWS_TRACE=1 PP_CONV_FORCE=rows PP_CONV_ROWS_PAIR=1 PP_LAYERS="dec5b 512,enc3b" timeout 100 python tests/bench_conv_layers.py 2>&1 | grep -E "trace" | sort -u | head -6
PP_CONV_TUNE_DEBUG=1 timeout 200 python tests/bench_conv_layers.py tuned > gpurun_out/r02_layers_tuned.txt 2> gpurun_out/r02_layers_tuned.err
cat gpurun_out/r02_layers_tuned.txt
