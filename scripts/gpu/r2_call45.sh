#!/bin/bash
mkdir -p gpurun_out
RMP_EDGE_TUNING=1,11 timeout 120 python -m pytest tests/test_tc_kernels.py -x -q -m gpu > gpurun_out/r2c45_tc.log 2>&1; tail -3 gpurun_out/r2c45_tc.log
timeout 120 python scripts/edge_variants.py 20 > gpurun_out/r2c45_edge_variants.log 2>&1; tail -8 gpurun_out/r2c45_edge_variants.log
