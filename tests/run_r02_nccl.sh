#!/bin/bash
# round 2: effect of the NCCL CTA budget on the 8-GPU step (all-reduce CTAs compete with the conv kernels for SMs)
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
for ctas in default 8 4; do
  if [ "$ctas" = "default" ]; then E=""; else E="NCCL_MAX_CTAS=$ctas"; fi
  env $E PP_NCCL_TEST=1 timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-same-box --no-other-bn --no-profile-pass 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('NCCL_MAX_CTAS=$ctas', round(d['value'],1), 'img/s', round(d['ms_per_step'],3), round(d['ms_per_step_median'],3), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],3))"
done
