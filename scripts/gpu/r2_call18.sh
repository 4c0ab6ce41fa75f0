#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -u -m pytest tests/test_x3_kernels.py -x -q > gpurun_out/r2c18_x3.log 2>&1; tail -3 gpurun_out/r2c18_x3.log
X3_TIMELINE=1 timeout 60 python scripts/x3_launch.py 5 30 16384 edge 2>&1 | head -12 | tee gpurun_out/r2c18_tl_edge.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision fp32x3 > gpurun_out/r2c18_bench_x3.json 2> gpurun_out/r2c18_bench_x3.err; tail -3 gpurun_out/r2c18_bench_x3.err
python -c "
import json; d=json.load(open('gpurun_out/r2c18_bench_x3.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d['roofline']['kernel_ms'], d['roofline']['kernel_share_of_step'])"
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 280 -c 120 --csv --log-file gpurun_out/r2c18_x3_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --precision fp32x3 --no-graph > gpurun_out/r2c18_ncu.log 2>&1
tail -1 gpurun_out/r2c18_ncu.log | cut -c1-200
