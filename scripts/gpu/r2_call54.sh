#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -u -m pytest tests/test_training_path.py -x -q > gpurun_out/r2c54_train.log 2>&1; tail -25 gpurun_out/r2c54_train.log
timeout 300 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/r2c54_bench_c4.json 2> gpurun_out/r2c54_err.log; tail -1 gpurun_out/r2c54_err.log
RMP_TRAIN_FUSED_FWD=0 timeout 300 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/r2c54_bench_c4_unfused.json 2> gpurun_out/r2c54_err2.log; tail -1 gpurun_out/r2c54_err2.log
python -c "
import json
for f in ('gpurun_out/r2c54_bench_c4.json','gpurun_out/r2c54_bench_c4_unfused.json'):
    d=json.loads([l for l in open(f) if l.startswith('{')][0]); print('c4', d['ms_per_step'], d['value'], d['loss'], d['gpu_launches'])"
