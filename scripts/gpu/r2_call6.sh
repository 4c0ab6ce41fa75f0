#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -u -m pytest tests/test_training_path.py -x -q -s > gpurun_out/r2c6_train.log 2>&1; echo "exit $?" >> gpurun_out/r2c6_train.log
tail -30 gpurun_out/r2c6_train.log | cut -c1-300
if grep -q "exit 0" gpurun_out/r2c6_train.log; then
  timeout 600 python bench.py --workload c4 --steps 10 --warmup 3 > gpurun_out/r2c6_bench_c4.json 2> gpurun_out/r2c6_bench_c4.err
  python -c "
import json; d=json.load(open('gpurun_out/r2c6_bench_c4.json')); print(d['ms_per_step'], d['value'], d['e2e'], d['gpu_launches'], d['config']['launch'], d['loss'])"
  tail -5 gpurun_out/r2c6_bench_c4.err
  timeout 600 python scripts/train_profile.py > gpurun_out/r2c7_train_profile.txt 2>&1
  grep -E "Name|rmp|void|Self CUDA time" gpurun_out/r2c7_train_profile.txt | cut -c1-75,150-230 | head -24
fi
