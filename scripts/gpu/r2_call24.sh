#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -u -m pytest tests/test_rnampnn_cuda.py tests/test_rnampnn_dropin.py tests/test_forest_head.py -x -q > gpurun_out/r2c24_rnampnn.log 2>&1; tail -5 gpurun_out/r2c24_rnampnn.log
for prec in bf16 fp32; do
  RMP_GRAPH_NORM_TWO_LAUNCH=1 timeout 300 python scripts/bench_rnampnn.py --precision $prec --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c24_rnampnn_${prec}_two.json 2> gpurun_out/r2c24_err.log; tail -2 gpurun_out/r2c24_err.log
  timeout 300 python scripts/bench_rnampnn.py --precision $prec --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c24_rnampnn_${prec}_fused.json 2> gpurun_out/r2c24_err.log; tail -2 gpurun_out/r2c24_err.log
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c24_rnampnn_*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); print(f, d['ms_per_step'], d['value'], d.get('gpu_launches'))
PY
