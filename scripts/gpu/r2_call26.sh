#!/bin/bash
# round-2 ncu evidence: launch lists of the bench command (bf16 default, fp32x3), --set full of the dominant kernels at C3 / fp32x3
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bf16_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph > gpurun_out/r2c26_ncu1.log 2>&1; tail -1 gpurun_out/r2c26_ncu1.log | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_fp32x3_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph --precision fp32x3 > gpurun_out/r2c26_ncu2.log 2>&1; tail -1 gpurun_out/r2c26_ncu2.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:edge_mlp_tc -s 3 -c 1 -f -o gpurun_out/r2_edge_c3_full python scripts/edge_launch.py 5 48 32768 > gpurun_out/r2c26_ncu3.log 2>&1; tail -1 gpurun_out/r2c26_ncu3.log
ncu --set full --clock-control none --import-source on -k regex:row_mlp_x3 -s 3 -c 1 -f -o gpurun_out/r2_x3_edge_full python scripts/x3_launch.py 5 30 16384 edge > gpurun_out/r2c26_ncu4.log 2>&1; tail -1 gpurun_out/r2c26_ncu4.log
ncu --set full --clock-control none --import-source on -k regex:edge_mlp_tc -s 3 -c 1 -f -o gpurun_out/r2_edge_c2_full python scripts/edge_launch.py 5 30 16384 > gpurun_out/r2c26_ncu5.log 2>&1; tail -1 gpurun_out/r2c26_ncu5.log
ls -la gpurun_out/*.ncu-rep | tail -5
timeout 300 python bench.py --workload c3 --no-cpu-baseline > gpurun_out/r2c26_bench_c3.json 2> gpurun_out/r2c26_bench_c3.err; cut -c1-300 gpurun_out/r2c26_bench_c3.json
