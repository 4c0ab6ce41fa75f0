#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 200 python -u -m pytest tests/test_tc_kernels.py -x -q -s > gpurun_out/r2c3_tc.log 2>&1; echo "tc exit $?" >> gpurun_out/r2c3_tc.log
tail -3 gpurun_out/r2c3_tc.log
if grep -q "tc exit 0" gpurun_out/r2c3_tc.log; then
  timeout 120 python scripts/edge_timeline.py > gpurun_out/r2c3_edge_tl.log 2>&1
  head -36 gpurun_out/r2c3_edge_tl.log
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c3_bench.json 2> gpurun_out/r2c3_bench.err
  python -c "
import json; d=json.load(open('gpurun_out/r2c3_bench.json')); r=d['roofline']
print('value %.3e ms %.3f frac %.3f kernel_ms %.4f'%(d['value'],d['ms_per_step'],r['frac'],r['kernel_ms']))"
fi
