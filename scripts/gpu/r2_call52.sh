#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/train_profile.py > gpurun_out/r2c52_train_profile.txt 2>&1; tail -3 gpurun_out/r2c52_train_profile.txt | cut -c1-100
