#!/bin/bash
mkdir -p gpurun_out
X3_TIMELINE=1 timeout 60 python scripts/x3_launch.py 5 30 16384 edge 2>&1 | head -20 | tee gpurun_out/r2c17_tl_edge.log
