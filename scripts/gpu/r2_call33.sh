#!/bin/bash
mkdir -p gpurun_out
timeout 40 python scripts/edge_timeline.py > gpurun_out/r2c33_tl.log 2>&1; sed -n '/hand-off chain/,$p' gpurun_out/r2c33_tl.log
