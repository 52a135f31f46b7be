#!/bin/bash
run() { env "$@" timeout 300 python bench.py --bn eval --steps 40 --warmup 6 --no-cpu-baseline --no-same-box --no-e2e --no-other-bn --no-profile-pass 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('eval $*', round(d['ms_per_step'],3), round(d['ms_per_step_median'],3))"; }
run A=1
run PP_FWD_PARTS=0
run PP_NO_PRIORITY=1
run PP_BN_EVAL_BPS=4
run PP_CONV_ROWS=0
run A=2
