#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -u -m pytest tests/test_rnampnn_cuda.py -x -q -k "attention" > gpurun_out/r2c28_attn.log 2>&1; tail -25 gpurun_out/r2c28_attn.log
timeout 60 python scripts/attn_bench.py 64 256 8 2>&1 | tail -1 | tee gpurun_out/r2c28_attn_bench.log
timeout 60 python scripts/attn_bench.py 8 4096 8 2>&1 | tail -1 | tee -a gpurun_out/r2c28_attn_bench.log
timeout 60 python scripts/attn_bench.py 64 256 4 2>&1 | tail -1 | tee -a gpurun_out/r2c28_attn_bench.log
timeout 600 python -u -m pytest tests/test_rnampnn_cuda.py tests/test_rnampnn_dropin.py -x -q > gpurun_out/r2c28_rnampnn.log 2>&1; tail -5 gpurun_out/r2c28_rnampnn.log
timeout 300 python scripts/bench_rnampnn.py --precision bf16 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c28_rnampnn_bf16.json 2> gpurun_out/r2c28_err.log; tail -2 gpurun_out/r2c28_err.log; cut -c1-200 gpurun_out/r2c28_rnampnn_bf16.json
timeout 300 python scripts/bench_rnampnn.py --precision bf16 --workload r3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c28_rnampnn_bf16_r3.json 2> gpurun_out/r2c28_err.log; tail -2 gpurun_out/r2c28_err.log; cut -c1-200 gpurun_out/r2c28_rnampnn_bf16_r3.json
