#!/bin/bash
# ping-pong schedules of the edge kernel: correctness of every mode, then launch time per mode
mkdir -p gpurun_out
for m in s d p w e; do
  echo "=== schedule $m" | tee -a gpurun_out/r2c21_sched.log
  RMP_EDGE_SCHEDULE=$m timeout 300 python -m pytest tests/test_tc_kernels.py -x -q -k "edge_mlp" 2>&1 | tail -2 | tee -a gpurun_out/r2c21_sched.log
  RMP_EDGE_SCHEDULE=$m timeout 60 python scripts/edge_launch.py 20 2>&1 | tail -1 | tee -a gpurun_out/r2c21_sched.log
  RMP_EDGE_SCHEDULE=$m timeout 60 python scripts/edge_launch.py 20 48 32768 2>&1 | tail -1 | tee -a gpurun_out/r2c21_sched.log
done
RMP_EDGE_SCHEDULE=w timeout 60 python scripts/edge_timeline.py > gpurun_out/r2c21_tl_w.log 2>&1; head -40 gpurun_out/r2c21_tl_w.log
