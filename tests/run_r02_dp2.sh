#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tests/run_dp_check.py > gpurun_out/r02_dp_check_${N}gpu.log 2>&1; echo "dp check exit $?"
grep -E "step|DP CHECK|Error|error" gpurun_out/r02_dp_check_${N}gpu.log | tail -8
timeout 600 $TR bench.py --gpus $N --steps 30 --warmup 6 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; echo "bench $N exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_${N}gpu.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','n_gpus','ms_per_step','ms_per_step_median','gpu_launches')}, d['e2e']['value'], d['e2e']['ms_per_step'], d['extra'])"
tail -3 gpurun_out/r02_bench_${N}gpu.err
