#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -u -m pytest tests/test_x3_kernels.py tests/test_baseline_configs.py tests/test_valid_set_recovery.py tests/test_cuda_parity.py -q -s > gpurun_out/r2c20_x3.log 2>&1; grep -E "fp32x3|passed|failed" gpurun_out/r2c20_x3.log | tail -12
