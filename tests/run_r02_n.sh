#!/bin/bash
run() { env "$@" timeout 300 python bench.py --steps 40 --warmup 6 --no-cpu-baseline --no-same-box --no-e2e --no-other-bn --no-profile-pass 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', round(d['ms_per_step'],3), round(d['ms_per_step_median'],3))"; }
run A=1
run PP_BN_RED_BPS=4
run PP_BN_RED_BPS=3
run PP_BN_RED_BPS=12
run PP_CONV_ROWS=0
run PP_CONV_ROWS_PAIR=0
run PP_ROWS_TMA_STORE=0
run A=2
