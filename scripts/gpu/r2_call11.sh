#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -u -m pytest tests/test_x3_kernels.py -x -q -s > gpurun_out/r2c11_x3.log 2>&1; tail -25 gpurun_out/r2c11_x3.log
