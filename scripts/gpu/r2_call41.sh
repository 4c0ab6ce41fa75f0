#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/edge_variants.py 20 > gpurun_out/r2c41_edge_variants.log 2>&1; tail -20 gpurun_out/r2c41_edge_variants.log
timeout 200 python scripts/train_profile.py > gpurun_out/r2c41_train_profile.txt 2>&1; tail -3 gpurun_out/r2c41_train_profile.txt | cut -c1-100
