#!/bin/bash
mkdir -p gpurun_out
python scripts/gemm_rows_bench.py > gpurun_out/r2c10_gemm_rows.log 2>&1; cat gpurun_out/r2c10_gemm_rows.log
timeout 600 python -u -m pytest tests/test_training_path.py -x -q > gpurun_out/r2c9_train.log 2>&1; tail -3 gpurun_out/r2c9_train.log
timeout 600 python bench.py --workload c4 --steps 20 --warmup 3 > gpurun_out/r2c9_bench_c4.json 2> gpurun_out/r2c9_bench_c4.err
python -c "
import json; d=json.load(open('gpurun_out/r2c9_bench_c4.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d['loss'])"
