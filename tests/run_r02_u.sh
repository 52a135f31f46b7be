#!/bin/bash
# round 2, session 3: low-resolution aux logits in the fused loss + replicated eval-BN backward accumulators
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "scribble_loss or golden or loss_functions or memory or graph or max_ch_728 or compact or bn_eval or first_conv_head" ) > gpurun_out/r02_u_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02_u_pytest.log
B="python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-same-box"
$B > gpurun_out/r02_u_bench.json 2> gpurun_out/r02_u_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_u_bench.json"))
print("train %.3f ms (median %.3f)  eval %.3f ms  roofline %.3f  e2e %.3f" % (
    d["ms_per_step"], d["ms_per_step_median"], d["extra"]["other_bn_regime"]["ms_per_step"], d["roofline"]["frac"],
    d["e2e"]["ms_per_step"]))
print(json.dumps(d["roofline"]["other_kernels"]["scribble_loss_fwd+bwd (HBM bound)"]))
PY
