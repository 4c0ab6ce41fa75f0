#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/edge_variants.py 20 > gpurun_out/r2c42_edge_variants.log 2>&1; tail -12 gpurun_out/r2c42_edge_variants.log
RMP_EDGE_TUNING=1,7 timeout 300 python -m pytest tests/test_tc_kernels.py -x -q -m gpu > gpurun_out/r2c42_tc.log 2>&1; tail -3 gpurun_out/r2c42_tc.log
