#!/bin/bash
mkdir -p gpurun_out
X3_TIMELINE=1 timeout 60 python scripts/x3_launch.py 5 30 16384 edge 2>&1 | tee gpurun_out/r2c15_tl_edge.log
X3_TIMELINE=1 timeout 60 python scripts/x3_launch.py 5 30 16384 noadds 2>&1 | tee gpurun_out/r2c15_tl_noadds.log
