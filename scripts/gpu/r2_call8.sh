#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --workload c4 --steps 10 --warmup 3 > gpurun_out/r2c8_bench_c4.json 2> gpurun_out/r2c8_bench_c4.err
python -c "
import json; d=json.load(open('gpurun_out/r2c8_bench_c4.json')); print(d['ms_per_step'], d['value'], d['e2e'], d['gpu_launches'], d['config']['launch'], d['loss'])"
grep -v "^\s*$" gpurun_out/r2c8_bench_c4.err | tail -12 | cut -c1-200
