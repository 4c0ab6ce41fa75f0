#!/bin/bash
mkdir -p gpurun_out
for v in edge noadds nogelu ffn; do python scripts/x3_launch.py 5 30 16384 $v; done 2>&1 | tee gpurun_out/r2c13_variants.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:row_mlp_x3 -s 2 -c 1 -o gpurun_out/r2c13_x3 -f python scripts/x3_launch.py 1 > gpurun_out/r2c13_ncu.log 2>&1
tail -2 gpurun_out/r2c13_ncu.log
