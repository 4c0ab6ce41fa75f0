#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/edge_train_fwd_bench.py > gpurun_out/r2c56_edge_train.log 2>&1; cat gpurun_out/r2c56_edge_train.log | tail -8
