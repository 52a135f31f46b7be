#!/bin/bash
# wgrad_reduce_unpack with four splits in flight: parity + step time
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "wgrad_oihw or full_tile or golden or graph" 2>&1 | grep -E "^FAILED|passed|failed"
B="python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-same-box --no-e2e"
for i in 1 2; do $B 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); o=d['roofline']['other_kernels']
print('train %.3f ms (median %.3f)  eval %.3f ms  conv frac %.3f  wgrad %.3f ms/step  clocks %s' % (d['ms_per_step'], d['ms_per_step_median'], d['extra']['other_bn_regime']['ms_per_step'], d['roofline']['frac'], o['conv3x3_wgrad_tc_kernel']['kernel_ms_per_step'], d['clocks']['sm_mhz']))"; done
