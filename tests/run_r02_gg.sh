#!/bin/bash
mkdir -p gpurun_out
timeout 170 python -m pytest -q "tests/test_gpu_fullsize.py::test_fullsize_trained_state_north_star[train-config4_upper_256_C5]" "tests/test_gpu_fullsize.py::test_fullsize_init_state_vs_oracle[config1_baseline_256_C5]" 2>&1 | tail -5
