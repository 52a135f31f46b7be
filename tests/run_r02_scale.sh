#!/bin/bash
# round 2, multi-GPU lines: usage run_r02_scale.sh N [quick]   (BASELINE.json configs 2-5 at N GPUs of one node)
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
S="--steps 30 --warmup 6 --no-cpu-baseline --no-same-box"
timeout 600 $TR tests/run_dp_check.py > gpurun_out/r02_dp_check_${N}gpu.log 2>&1; echo "dp check exit $?"; grep -E "DP CHECK" gpurun_out/r02_dp_check_${N}gpu.log
timeout 600 $TR bench.py --gpus $N $S > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; echo "config2 exit $?"
timeout 600 $TR bench.py --gpus $N $S --classes 4 > gpurun_out/r02_bench_${N}gpu_config3_acdc256.json 2> /dev/null; echo "config3 (256) exit $?"
if [ "$2" != "quick" ]; then

timeout 600 $TR bench.py --gpus $N $S --workload upperbound > gpurun_out/r02_bench_${N}gpu_config4_upperbound.json 2> /dev/null; echo "config4 exit $?"
for B in 12 96; do
timeout 600 $TR bench.py --gpus $N --steps 15 --warmup 4 --no-cpu-baseline --no-same-box --no-e2e --no-other-bn --classes 2 --size 224 --batch $B > gpurun_out/r02_bench_${N}gpu_config5_lvsc224_b$B.json 2> /dev/null; echo "config5 b=$B exit $?"
done
fi
for f in gpurun_out/r02_bench_${N}gpu*.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], round(d['value'],1), 'img/s', round(d['ms_per_step'],3), 'ms', 'e2e', d['e2e'] and round(d['e2e']['value'],1))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
