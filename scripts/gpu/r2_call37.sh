#!/bin/bash
mkdir -p gpurun_out
for m in 20 0 5 100; do
  echo "=== spin $m" | tee -a gpurun_out/r2c37.log
  RMP_EDGE_SPIN=$m timeout 25 python scripts/edge_launch.py 20 2>&1 | tail -1 | tee -a gpurun_out/r2c37.log
  RMP_EDGE_SPIN=$m timeout 25 python scripts/edge_launch.py 20 48 32768 2>&1 | tail -1 | tee -a gpurun_out/r2c37.log
done
RMP_EDGE_SPIN=0 timeout 40 python scripts/edge_timeline.py > gpurun_out/r2c37_tl.log 2>&1; sed -n "/hand-off chain/,\$p" gpurun_out/r2c37_tl.log | head -8
