#!/bin/bash
# session-3 ncu evidence for the shipped edge kernel (variant 1, flags 3, elect.sync issuers): launch list of the bench command, --set full at C2 / C3
mkdir -p gpurun_out
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/r2c49_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:rmp:: -c 300 --csv --log-file gpurun_out/r2s3_bf16_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/r2c49_ncu1.log 2>&1; tail -1 gpurun_out/r2c49_ncu1.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:edge_mlp_tc -s 3 -c 1 -f -o gpurun_out/r2s3_edge_c2_full python scripts/edge_launch.py 5 30 16384 > gpurun_out/r2c49_ncu2.log 2>&1; tail -1 gpurun_out/r2c49_ncu2.log
ncu --set full --clock-control none --import-source on -k regex:edge_mlp_tc -s 3 -c 1 -f -o gpurun_out/r2s3_edge_c3_full python scripts/edge_launch.py 5 48 32768 > gpurun_out/r2c49_ncu3.log 2>&1; tail -1 gpurun_out/r2c49_ncu3.log
ls -la gpurun_out/r2s3*
