#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/train_profile.py > gpurun_out/r2c39_train_profile.txt 2>&1; grep -E "Self CUDA time total|^rmp|^void|nvjet|Name" gpurun_out/r2c39_train_profile.txt | cut -c1-150 | head -50
