#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/edge_variants.py 20 > gpurun_out/r2c44_edge_variants.log 2>&1; tail -8 gpurun_out/r2c44_edge_variants.log
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2c44_pytest.log 2>&1; tail -3 gpurun_out/r2c44_pytest.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c44_bench_c2.json 2> gpurun_out/r2c44_err.log; tail -2 gpurun_out/r2c44_err.log
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2c44_bench_c2.json') if l.startswith('{')][0]); print('c2', d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'])"
RMP_EDGE_TUNING=1,3 timeout 200 python scripts/edge_timeline.py > gpurun_out/r2c44_edge_tl.log 2>&1; tail -20 gpurun_out/r2c44_edge_tl.log
