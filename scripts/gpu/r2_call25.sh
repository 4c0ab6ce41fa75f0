#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -u -m pytest tests -m gpu -x -q > gpurun_out/r2c25_pytest.log 2>&1; tail -4 gpurun_out/r2c25_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/r2c25_smoke.log
timeout 600 python bench.py > gpurun_out/r2c25_bench.json 2> gpurun_out/r2c25_bench.err; tail -2 gpurun_out/r2c25_bench.err; cut -c1-600 gpurun_out/r2c25_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c25_bench_ref.json 2> gpurun_out/r2c25_bench_ref.err; cut -c1-400 gpurun_out/r2c25_bench_ref.json
