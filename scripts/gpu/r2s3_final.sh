#!/bin/bash
# final verification of round 2 (session 3): full GPU suite, smoke, plain run of the every-kernel-family script, bench lines of every
# single-GPU workload / arm, reference arm
mkdir -p gpurun_out
timeout 900 python -u -m pytest tests -m gpu -x -q > gpurun_out/r2s3f_pytest.log 2>&1; tail -3 gpurun_out/r2s3f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/r2s3f_smoke.log
timeout 300 python scripts/sanitize_smoke.py > gpurun_out/r2s3f_allkernels_plain.log 2>&1; tail -2 gpurun_out/r2s3f_allkernels_plain.log
timeout 600 python bench.py > gpurun_out/r2s3f_bench_bf16_c2.json 2> gpurun_out/r2s3f_err.log; tail -1 gpurun_out/r2s3f_err.log
timeout 300 python bench.py --precision fp32x3 --no-cpu-baseline > gpurun_out/r2s3f_bench_fp32x3_c2.json 2> gpurun_out/r2s3f_err.log; tail -1 gpurun_out/r2s3f_err.log
timeout 300 python bench.py --workload c3 --no-cpu-baseline > gpurun_out/r2s3f_bench_bf16_c3.json 2> gpurun_out/r2s3f_err.log; tail -1 gpurun_out/r2s3f_err.log
timeout 300 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/r2s3f_bench_c4_1gpu.json 2> gpurun_out/r2s3f_err.log; tail -1 gpurun_out/r2s3f_err.log
timeout 300 python scripts/bench_rnampnn.py > gpurun_out/r2s3f_rnampnn.log 2>&1; grep "^{" gpurun_out/r2s3f_rnampnn.log | tail -1 > gpurun_out/r2s3f_bench_rnampnn_bf16.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2s3f_bench_reference.json 2> gpurun_out/r2s3f_err.log; tail -1 gpurun_out/r2s3f_err.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2s3f_bench_*.json')):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); r=d.get('roofline') or {}
            print(f.split('/')[-1], 'ms', round(d.get('ms_per_step',0),3), 'value', round(d.get('value',0)), 'e2e', round((d.get('e2e') or {}).get('value',0)), 'frac', r.get('frac'), 'traffic', r.get('traffic'), 'launches', d.get('gpu_launches'), (d.get('cpu_baseline') or {}).get('kind'))
PY
