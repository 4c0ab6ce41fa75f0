#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --precision fp32x3 --no-cpu-baseline > gpurun_out/r2c46_bench_fp32x3_c2.json 2> gpurun_out/r2c46_err.log; tail -1 gpurun_out/r2c46_err.log
timeout 300 python bench.py --workload c3 --no-cpu-baseline > gpurun_out/r2c46_bench_bf16_c3.json 2> gpurun_out/r2c46_err.log; tail -1 gpurun_out/r2c46_err.log
timeout 300 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/r2c46_bench_c4_1gpu.json 2> gpurun_out/r2c46_err.log; tail -1 gpurun_out/r2c46_err.log
timeout 300 python scripts/bench_rnampnn.py > gpurun_out/r2c46_rnampnn.log 2>&1; tail -4 gpurun_out/r2c46_rnampnn.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c46_bench_*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); r=d.get('roofline') or {}
            print(f.split('/')[-1], 'ms', round(d.get('ms_per_step',0),3), 'value', round(d.get('value',0)), 'e2e', round((d.get('e2e') or {}).get('value',0)), 'frac', r.get('frac'), 'kernel_ms', r.get('kernel_ms'))
PY
