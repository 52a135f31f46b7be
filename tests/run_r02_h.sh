#!/bin/bash
mkdir -p gpurun_out
export PP_CONV_FORCE=rows PP_CONV_ROWS_PAIR=1 PP_LAYERS="dec5b 512,enc3b"
python tests/bench_conv_layers.py pair > gpurun_out/r02_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv3x3_rows -s 6 -c 2 -f -o gpurun_out/r02_prof_rows_pair python tests/bench_conv_layers.py pair > gpurun_out/r02_prof_ncu.log 2>&1
echo "ncu rc=$?"; cat gpurun_out/r02_prof_plain.log
