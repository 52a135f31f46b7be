#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x ) > gpurun_out/r02_pytest_o.log 2>&1
echo "parity tests rc=$?"; tail -2 gpurun_out/r02_pytest_o.log
PP_GRAPHS_DEBUG=1 timeout 300 python bench.py --steps 12 --warmup 4 --no-cpu-baseline --no-same-box --no-e2e --no-other-bn --no-profile-pass 2>&1 >/dev/null | grep "pp graph" | sort | uniq -c
run() { env "$@" timeout 300 python bench.py --steps 40 --warmup 6 --no-cpu-baseline --no-same-box --no-other-bn --no-profile-pass 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('$*', 'dev', round(d['ms_per_step'],3), round(d['ms_per_step_median'],3), 'e2e', round(e['ms_per_step'],3), 'item', round(e['ms_per_step_blocking_item_reads'],3), 'compact', round(e['compact_input']['ms_per_step'],3), 'launches', d['gpu_launches'])"; }
run A=1
run PP_GRAPHS=0
run A=2
run PP_GRAPHS=0 B=2
