#!/bin/bash
mkdir -p gpurun_out
for m in p o; do
  echo "=== issuer $m" | tee -a gpurun_out/r2c35_issuer.log
  RMP_EDGE_ISSUER=$m timeout 25 python scripts/edge_launch.py 20 2>&1 | tail -1 | tee -a gpurun_out/r2c35_issuer.log
  if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "issuer $m failed / hung"; continue; fi
  RMP_EDGE_ISSUER=$m timeout 25 python scripts/edge_launch.py 20 48 32768 2>&1 | tail -1 | tee -a gpurun_out/r2c35_issuer.log
  RMP_EDGE_ISSUER=$m timeout 25 python scripts/edge_launch.py 20 30 100 2>&1 | tail -1 | tee -a gpurun_out/r2c35_issuer.log
  RMP_EDGE_ISSUER=$m timeout 120 python -m pytest tests/test_tc_kernels.py -x -q -k "edge_mlp or forward_bf16" 2>&1 | tail -2 | tee -a gpurun_out/r2c35_issuer.log
done
RMP_EDGE_ISSUER=o timeout 40 python scripts/edge_timeline.py > gpurun_out/r2c35_tl_o.log 2>&1; sed -n "/hand-off chain/,\$p" gpurun_out/r2c35_tl_o.log | head -12
