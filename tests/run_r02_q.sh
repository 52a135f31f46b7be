#!/bin/bash
PP_CONV_TUNE_DEBUG=1 PP_LAYERS="dec2a" timeout 100 python tests/bench_conv_layers.py 2>&1 | tail -12
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "conv3x3_forward or full_tile" 2>&1 | tail -2
run() { env "$@" timeout 300 python bench.py --steps 40 --warmup 6 --no-cpu-baseline --no-same-box --no-e2e --no-profile-pass 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', round(d['ms_per_step'],3), round(d['ms_per_step_median'],3), 'eval', round(d['extra']['other_bn_regime']['ms_per_step'],3))"; }
run A=1
run A=2
