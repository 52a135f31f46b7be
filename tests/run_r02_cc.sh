#!/bin/bash
# single-GPU repeatability of the DP check's two passes under switches
run() { echo "== $*"; env "$@" timeout 200 python tests/repeat_check.py 2>&1 | grep -E "^step|DP CHECK|Error|error" | head -8; }
run A=1
run PP_WGRAD_ROWS=1
run PP_GRAPHS=0
run PP_CONV_AUTOTUNE=0
