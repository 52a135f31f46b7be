#!/bin/bash
# narrow weight gradient with the horizontal taps packed into N (PP_WGRAD_ROWS=2) vs the row kernel (=1)
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_parity.py -m gpu -q -k "conv3x3_forward_dgrad_wgrad or full_tile" ) > gpurun_out/r02_v_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "^FAILED|passed|failed" gpurun_out/r02_v_pytest.log | head -20
B="python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-same-box --no-e2e"
for r in 2 1; do
  PP_WGRAD_ROWS=$r $B > gpurun_out/r02_v_bench_rows$r.json 2> gpurun_out/r02_v_bench_rows$r.err; echo "bench rows=$r rc=$?"
done
python - <<'PY'
import json
for r in (2, 1):
    try:
        d = json.load(open("gpurun_out/r02_v_bench_rows%d.json" % r))
        o = d["roofline"]["other_kernels"]
        print("rows=%d train %.3f ms (median %.3f)  eval %.3f ms  conv frac %.3f  wgrad %.3f ms/step frac %.3f" % (
            r, d["ms_per_step"], d["ms_per_step_median"], d["extra"]["other_bn_regime"]["ms_per_step"], d["roofline"]["frac"],
            o["conv3x3_wgrad_tc_kernel"]["kernel_ms_per_step"], o["conv3x3_wgrad_tc_kernel"]["frac"]))
    except Exception as e:
        print(r, "failed", e)
PY
