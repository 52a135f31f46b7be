#!/bin/bash
# 128..256-channel sources of the narrow layers through the kx-in-N row kernel (64-channel column groups)
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_parity.py -m gpu -q -k "conv3x3_forward_dgrad_wgrad or full_tile or wgrad_oihw or golden" ) > gpurun_out/r02_w_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "^FAILED|passed|failed" gpurun_out/r02_w_pytest.log | head -20
B="python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-same-box --no-e2e"
$B > gpurun_out/r02_w_bench.json 2> gpurun_out/r02_w_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_w_bench.json"))
o = d["roofline"]["other_kernels"]
print("train %.3f ms (median %.3f)  eval %.3f ms  conv frac %.3f  wgrad %.3f ms/step frac %.3f" % (
    d["ms_per_step"], d["ms_per_step_median"], d["extra"]["other_bn_regime"]["ms_per_step"], d["roofline"]["frac"],
    o["conv3x3_wgrad_tc_kernel"]["kernel_ms_per_step"], o["conv3x3_wgrad_tc_kernel"]["frac"]))
PY
