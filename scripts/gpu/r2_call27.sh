#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:rmp:: -c 300 --csv --log-file gpurun_out/r2_bf16_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/r2c27_ncu1.log 2>&1; tail -1 gpurun_out/r2c27_ncu1.log | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:rmp:: -c 300 --csv --log-file gpurun_out/r2_fp32x3_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph --precision fp32x3 > gpurun_out/r2c27_ncu2.log 2>&1; tail -1 gpurun_out/r2c27_ncu2.log | cut -c1-200
wc -l gpurun_out/r2_bf16_launches.csv gpurun_out/r2_fp32x3_launches.csv
