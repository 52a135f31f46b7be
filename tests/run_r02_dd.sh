#!/bin/bash
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
run() { echo "== $*"; env "$@" timeout 200 $TR tests/run_dp_check.py 2>&1 | grep -E "^step|^    |DP CHECK|Error" | head -14; }
run PP_DP_DIAG=1
run PP_DP_DIAG=0 PP_WGRAD_ROWS=1
run PP_DP_DIAG=0 PP_GRAPHS=0
