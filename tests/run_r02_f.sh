#!/bin/bash
mkdir -p gpurun_out
for mode in 0 1; do
  ( PP_CONV_ROWS_PAIR=$mode timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "conv3x3_forward or full_tile or bn_eval_fused" ) > gpurun_out/r02_pytest_f_pair$mode.log 2>&1
  echo "pair=$mode op tests rc=$?"; tail -2 gpurun_out/r02_pytest_f_pair$mode.log
  PP_CONV_ROWS_PAIR=$mode PP_CONV_ROWS_DEBUG=1 timeout 200 python tests/bench_conv_layers.py pair$mode > gpurun_out/r02_layers_pair$mode.txt 2> gpurun_out/r02_layers_pair$mode.err
  echo "layers rc=$?"
done
paste gpurun_out/r02_layers_pair0.txt gpurun_out/r02_layers_pair1.txt | cut -c1-150
