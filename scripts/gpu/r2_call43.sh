#!/bin/bash
mkdir -p gpurun_out
RMP_EDGE_TUNING=1,3 timeout 200 python scripts/edge_timeline.py > gpurun_out/r2c43_edge_tl.log 2>&1; tail -25 gpurun_out/r2c43_edge_tl.log
