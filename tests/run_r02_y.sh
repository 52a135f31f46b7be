#!/bin/bash
# kx-in-N row weight gradient with R-row strips: parity, kernels alone, step A/B against the round-1 row kernel
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_parity.py -m gpu -q -k "conv3x3_forward_dgrad_wgrad or full_tile or wgrad_oihw" ) > gpurun_out/r02_y_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "^FAILED|passed|failed" gpurun_out/r02_y_pytest.log | head -20
PP_WGRAD_ROWS=2 python tests/bench_wgrad_narrow.py > gpurun_out/r02_y_wgrad_narrow_alone.txt 2>&1; cat gpurun_out/r02_y_wgrad_narrow_alone.txt
B="python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-same-box --no-e2e"
for r in 2 1 2; do
  PP_WGRAD_ROWS=$r $B 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); o=d['roofline']['other_kernels']
print('rows=$r train %.3f ms (median %.3f)  eval %.3f ms  conv frac %.3f  wgrad %.3f ms/step' % (d['ms_per_step'], d['ms_per_step_median'], d['extra']['other_bn_regime']['ms_per_step'], d['roofline']['frac'], o['conv3x3_wgrad_tc_kernel']['kernel_ms_per_step']))"
done
