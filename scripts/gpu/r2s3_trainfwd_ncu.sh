#!/bin/bash
# ncu --set full of the fused training forward (edge_mlp_tc_kernel<0,0,5>, p_drop 0.1) at C4 shapes
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:edge_mlp_tc --launch-skip 24 -c 1 -f -o gpurun_out/r2s3_edge_trainfwd_full python scripts/edge_train_fwd_bench.py > gpurun_out/r2s3_trainfwd_ncu.log 2>&1; tail -2 gpurun_out/r2s3_trainfwd_ncu.log
ls -la gpurun_out/r2s3_edge_trainfwd_full.ncu-rep
