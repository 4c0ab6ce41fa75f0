#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -u -m pytest tests/test_tc_kernels.py tests/test_training_path.py -x -q -m gpu -k "edge or tc or forward or native_training_step_grad" > gpurun_out/r2c63_tests.log 2>&1; tail -3 gpurun_out/r2c63_tests.log
timeout 60 python scripts/edge_variants.py 20 > gpurun_out/r2c63_edge_variants.log 2>&1; tail -2 gpurun_out/r2c63_edge_variants.log
timeout 120 python bench.py --no-cpu-baseline > gpurun_out/r2c63_bench_c2.json 2> gpurun_out/r2c63_err.log; tail -1 gpurun_out/r2c63_err.log
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2c63_bench_c2.json') if l.startswith('{')][0]); print('c2', d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'])"
