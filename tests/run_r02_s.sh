#!/bin/bash
# round 2, session 3: replicated BatchNorm-backward accumulators + weight packing on the side stream.
# BatchNorm / golden / graph tests, then A/B bench lines (new default vs PP_BN_REPLICAS=1 PP_PACK_SIDE=0).
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bn or golden or graph or full_size_step or data_parallel or flat_adam" ) > gpurun_out/r02_s_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02_s_pytest.log
B="python bench.py --steps 30 --warmup 8 --no-cpu-baseline --no-same-box"
$B > gpurun_out/r02_s_bench_new.json 2> gpurun_out/r02_s_bench_new.err; echo "new rc=$?"
PP_BN_REPLICAS=1 PP_PACK_SIDE=0 $B --no-e2e > gpurun_out/r02_s_bench_old.json 2> /dev/null; echo "old rc=$?"
PP_PACK_SIDE=0 $B --no-e2e > gpurun_out/r02_s_bench_nopack.json 2> /dev/null; echo "nopack rc=$?"
python - <<'PY'
import json
for n in ("new", "old", "nopack"):
    try:
        d = json.load(open("gpurun_out/r02_s_bench_%s.json" % n))
        print(n, "train %.3f ms (median %.3f)  eval %.3f ms  roofline %.3f  e2e %s" % (
            d["ms_per_step"], d["ms_per_step_median"], d["extra"]["other_bn_regime"]["ms_per_step"], d["roofline"]["frac"],
            d.get("e2e", {}).get("ms_per_step")))
    except Exception as e:
        print(n, "failed", e)
PY
