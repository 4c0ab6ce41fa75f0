#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -u -m pytest tests/test_training_path.py -x -q -k "edge_mlp_train or native or dropin or decreases" > gpurun_out/r2c62_train.log 2>&1; tail -4 gpurun_out/r2c62_train.log
timeout 200 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/r2c62_bench_c4.json 2> gpurun_out/r2c62_err.log; tail -1 gpurun_out/r2c62_err.log
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2c62_bench_c4.json') if l.startswith('{')][0]); print('c4', d['ms_per_step'], d['value'], d['loss'], d['gpu_launches'])"
