#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/train_profile.py > gpurun_out/r2c7_train_profile.txt 2>&1
grep -E "Name|cuda|rmp|void|at::|Memcpy|Self CUDA time" gpurun_out/r2c7_train_profile.txt | cut -c1-220 | head -70
