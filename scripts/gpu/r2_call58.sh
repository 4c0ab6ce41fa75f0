#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -u -m pytest tests/test_training_path.py tests/test_dropin_modules.py tests/test_reference_scripts.py -x -q -m gpu > gpurun_out/r2c58_train.log 2>&1; tail -25 gpurun_out/r2c58_train.log
timeout 300 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/r2c58_bench_c4.json 2> gpurun_out/r2c58_err.log; tail -1 gpurun_out/r2c58_err.log
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2c58_bench_c4.json') if l.startswith('{')][0]); print('c4', d['ms_per_step'], d['value'], d['loss'], d['gpu_launches'])"
