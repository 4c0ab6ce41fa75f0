#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision fp32x3 > gpurun_out/r2c12_bench_x3.json 2> gpurun_out/r2c12_bench_x3.err; tail -3 gpurun_out/r2c12_bench_x3.err
python -c "
import json; d=json.load(open('gpurun_out/r2c12_bench_x3.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d['roofline']['kernel_ms'], d['roofline']['kernel_share_of_step'])"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision fp32 > gpurun_out/r2c12_bench_fp32.json 2> gpurun_out/r2c12_bench_fp32.err
python -c "
import json; d=json.load(open('gpurun_out/r2c12_bench_fp32.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d['roofline']['kernel_ms'], d['roofline']['kernel_share_of_step'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2c12_x3_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --precision fp32x3 --no-graph > gpurun_out/r2c12_ncu.log 2>&1
tail -2 gpurun_out/r2c12_ncu.log
