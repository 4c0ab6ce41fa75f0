#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -u -m pytest tests/test_x3_kernels.py tests/test_cuda_parity.py -x -q > gpurun_out/r2c19_x3.log 2>&1; tail -5 gpurun_out/r2c19_x3.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision fp32x3 > gpurun_out/r2c19_bench_x3.json 2> gpurun_out/r2c19_bench_x3.err; tail -3 gpurun_out/r2c19_bench_x3.err
python -c "
import json; d=json.load(open('gpurun_out/r2c19_bench_x3.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d['roofline']['kernel_ms'], d['roofline']['kernel_share_of_step'])"
