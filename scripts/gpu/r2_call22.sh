#!/bin/bash
# staggered schedule with warp-elected arrives (a), and with 1/2 (h) or 1/4 (q) of the GELU pairs on the FMA pipe
mkdir -p gpurun_out
for m in s a h q; do
  echo "=== schedule $m" | tee -a gpurun_out/r2c22_sched.log
  RMP_EDGE_SCHEDULE=$m timeout 300 python -m pytest tests/test_tc_kernels.py -x -q -k "edge_mlp or forward_bf16" 2>&1 | tail -3 | tee -a gpurun_out/r2c22_sched.log
  RMP_EDGE_SCHEDULE=$m timeout 60 python scripts/edge_launch.py 20 2>&1 | tail -1 | tee -a gpurun_out/r2c22_sched.log
  RMP_EDGE_SCHEDULE=$m timeout 60 python scripts/edge_launch.py 20 48 32768 2>&1 | tail -1 | tee -a gpurun_out/r2c22_sched.log
done
RMP_EDGE_SCHEDULE=h timeout 60 python scripts/edge_timeline.py > gpurun_out/r2c22_tl_h.log 2>&1; head -36 gpurun_out/r2c22_tl_h.log
