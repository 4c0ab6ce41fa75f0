#!/bin/bash
# round 2, GPU call 2: new edge kernel (two streams, TMEM gather operand, selector-added Pc): parity, bench, timeline
set -x
mkdir -p gpurun_out
timeout 150 python -u -m pytest tests/test_tc_kernels.py -x -v > gpurun_out/r2c2_tc.log 2>&1; echo "tc exit $?" >> gpurun_out/r2c2_tc.log
tail -5 gpurun_out/r2c2_tc.log
if grep -q "tc exit 0" gpurun_out/r2c2_tc.log; then
  timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c2_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c2_pytest.log
  tail -8 gpurun_out/r2c2_pytest.log
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c2_bench.json 2> gpurun_out/r2c2_bench.err
  cat gpurun_out/r2c2_bench.json
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload c3 > gpurun_out/r2c2_bench_c3.json 2>> gpurun_out/r2c2_bench.err
  cat gpurun_out/r2c2_bench_c3.json
  timeout 120 python scripts/edge_timeline.py > gpurun_out/r2c2_edge_tl.log 2>&1
  head -40 gpurun_out/r2c2_edge_tl.log
fi
