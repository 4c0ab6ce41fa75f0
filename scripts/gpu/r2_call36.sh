#!/bin/bash
mkdir -p gpurun_out
for m in s d; do
  echo "=== schedule $m" | tee -a gpurun_out/r2c36.log
  RMP_EDGE_SCHEDULE=$m timeout 25 python scripts/edge_launch.py 20 2>&1 | tail -1 | tee -a gpurun_out/r2c36.log
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "schedule $m failed / hung"; continue; fi
  RMP_EDGE_SCHEDULE=$m timeout 25 python scripts/edge_launch.py 20 48 32768 2>&1 | tail -1 | tee -a gpurun_out/r2c36.log
  RMP_EDGE_SCHEDULE=$m timeout 120 python -m pytest tests/test_tc_kernels.py -x -q -k "edge_mlp or forward_bf16" 2>&1 | tail -2 | tee -a gpurun_out/r2c36.log
done
timeout 40 python scripts/edge_timeline.py > gpurun_out/r2c36_tl.log 2>&1; sed -n "/hand-off chain/,\$p" gpurun_out/r2c36_tl.log | head -10
