#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x ) > gpurun_out/r02_pytest_m.log 2>&1
echo "parity tests rc=$?"; tail -2 gpurun_out/r02_pytest_m.log
for bps in 6 3 12; do
PP_BN_EVAL_BPS=$bps timeout 300 python bench.py --bn eval --steps 30 --warmup 5 --no-cpu-baseline --no-same-box --no-e2e --no-other-bn 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('eval bps=$bps', round(d['ms_per_step'],3), round(d['ms_per_step_median'],3))"
done
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-same-box --no-e2e > gpurun_out/r02_bench_m.json 2>gpurun_out/r02_bench_m.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_m.json').read().strip().splitlines()[-1]); print('train', round(d['ms_per_step'],3), round(d['ms_per_step_median'],3), d['extra'])"
