#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -u -m pytest tests/test_training_path.py -x -q > gpurun_out/r2c9_train.log 2>&1; tail -3 gpurun_out/r2c9_train.log
timeout 600 python bench.py --workload c4 --steps 20 --warmup 3 > gpurun_out/r2c9_bench_c4.json 2> gpurun_out/r2c9_bench_c4.err
python -c "
import json; d=json.load(open('gpurun_out/r2c9_bench_c4.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d['loss'])"
timeout 600 python scripts/train_profile.py > gpurun_out/r2c9_train_profile.txt 2>&1
grep -E "Name|rmp|void|Self CUDA time" gpurun_out/r2c9_train_profile.txt | cut -c1-75,150-230 | head -16
