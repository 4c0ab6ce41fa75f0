#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -u -m pytest tests/test_training_path.py -x -q -s > gpurun_out/r2c6_train.log 2>&1; echo "exit $?" >> gpurun_out/r2c6_train.log
tail -40 gpurun_out/r2c6_train.log
if grep -q "exit 0" gpurun_out/r2c6_train.log; then
  timeout 600 python bench.py --workload c4 --steps 10 --warmup 3 > gpurun_out/r2c6_bench_c4.json 2> gpurun_out/r2c6_bench_c4.err; cat gpurun_out/r2c6_bench_c4.json; tail -5 gpurun_out/r2c6_bench_c4.err
  timeout 600 python scripts/train_step_bench.py --steps 5 > gpurun_out/r2c6_interim.json 2>&1; tail -2 gpurun_out/r2c6_interim.json
fi
