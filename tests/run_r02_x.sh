#!/bin/bash
# narrow weight gradient alone (row kernel vs kx-in-N kernel) + one ncu --set full capture of the new kernel
mkdir -p gpurun_out
PP_WGRAD_ROWS=1 python tests/bench_wgrad_narrow.py > gpurun_out/r02_wgrad_narrow_alone.txt 2>&1
PP_WGRAD_ROWS=2 python tests/bench_wgrad_narrow.py >> gpurun_out/r02_wgrad_narrow_alone.txt 2>&1
cat gpurun_out/r02_wgrad_narrow_alone.txt
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"conv3x3_wgrad_rowsn_tc_kernel<.int.32, .int.32>" -s 3 -c 1 \
  -o gpurun_out/r02_prof_wgrad_rowsn32 -f python tests/bench_wgrad_narrow.py > gpurun_out/ncu_r02_prof_wgrad_rowsn32.log 2>&1
echo "capture exit $?"
ncu -i gpurun_out/r02_prof_wgrad_rowsn32.ncu-rep --page raw --csv > gpurun_out/r02_ncu_prof_wgrad_rowsn32_raw.csv 2>/dev/null
wc -c gpurun_out/r02_ncu_prof_wgrad_rowsn32_raw.csv
