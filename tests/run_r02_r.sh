#!/bin/bash
for i in 0 1 2 3 4 5; do
  echo "== plan $i"; PP_CONV_FORCE=rows PP_CONV_FORCE_PLAN=$i timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "conv3x3_forward or full_tile or bn_eval_fused" 2>&1 | grep -E "^FAILED|passed|failed" | head -8
done
