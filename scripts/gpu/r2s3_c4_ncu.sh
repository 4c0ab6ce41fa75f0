#!/bin/bash
# ncu launch list of about one C4 training step (all kernels: the library's and torch's), eager launches
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled --launch-skip 3000 -c 1150 --csv --log-file gpurun_out/r2s3_c4_launches.csv python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/r2s3_c4_ncu.log 2>&1; tail -1 gpurun_out/r2s3_c4_ncu.log | cut -c1-160
wc -l gpurun_out/r2s3_c4_launches.csv
