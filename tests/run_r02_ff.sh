#!/bin/bash
# final state of the round: full GPU suite + smoke + the headline bench line
mkdir -p gpurun_out; rm -f gpurun_out/r02_parity_fullsize.txt
( time python -m pytest tests/ -m gpu -q ) > gpurun_out/r02_pytest_gpu_final.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_gpu_final.log
python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke.log
python bench.py --no-cpu-baseline --no-same-box > gpurun_out/r02_bench_1gpu_last.json 2> /dev/null; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_1gpu_last.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['extra']['other_bn_regime']['ms_per_step'], d['clocks'])"
grep -c MISS gpurun_out/r02_parity_fullsize.txt
