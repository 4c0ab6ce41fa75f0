#!/bin/bash
mkdir -p gpurun_out
for m in l h; do
  echo "=== roles $m" | tee -a gpurun_out/r2c34_roles.log
  RMP_EDGE_ROLES=$m timeout 25 python scripts/edge_launch.py 20 2>&1 | tail -1 | tee -a gpurun_out/r2c34_roles.log
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "roles $m failed / hung"; continue; fi
  RMP_EDGE_ROLES=$m timeout 25 python scripts/edge_launch.py 20 48 32768 2>&1 | tail -1 | tee -a gpurun_out/r2c34_roles.log
  RMP_EDGE_ROLES=$m timeout 120 python -m pytest tests/test_tc_kernels.py -x -q -k "edge_mlp or forward_bf16" 2>&1 | tail -2 | tee -a gpurun_out/r2c34_roles.log
done
RMP_EDGE_ROLES=h timeout 40 python scripts/edge_timeline.py > gpurun_out/r2c34_tl_h.log 2>&1; sed -n "/hand-off chain/,\$p" gpurun_out/r2c34_tl_h.log | head -12
