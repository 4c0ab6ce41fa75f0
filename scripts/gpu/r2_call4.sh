#!/bin/bash
set -x
mkdir -p gpurun_out
python scripts/edge_launch.py 5 > gpurun_out/r2c4_edge_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:edge_mlp_tc -s 2 -c 1 -f -o gpurun_out/r2c4_edge_v2 \
    python scripts/edge_launch.py 5 > gpurun_out/r2c4_ncu.log 2>&1
cat gpurun_out/r2c4_edge_plain.log
