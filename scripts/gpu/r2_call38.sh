#!/bin/bash
mkdir -p gpurun_out
for m in 0 1; do
  echo "=== phase_sync $m" | tee -a gpurun_out/r2c38.log
  RMP_EDGE_PHASE_SYNC=$m timeout 25 python scripts/edge_launch.py 20 2>&1 | tail -1 | tee -a gpurun_out/r2c38.log
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "phase_sync $m failed / hung"; continue; fi
  RMP_EDGE_PHASE_SYNC=$m timeout 25 python scripts/edge_launch.py 20 48 32768 2>&1 | tail -1 | tee -a gpurun_out/r2c38.log
  RMP_EDGE_PHASE_SYNC=$m timeout 25 python scripts/edge_launch.py 20 30 100 2>&1 | tail -1 | tee -a gpurun_out/r2c38.log
  RMP_EDGE_PHASE_SYNC=$m timeout 120 python -m pytest tests/test_tc_kernels.py -x -q -k "edge_mlp or forward_bf16" 2>&1 | tail -2 | tee -a gpurun_out/r2c38.log
done
