#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2c5_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c5_pytest.log
tail -6 gpurun_out/r2c5_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c5_bench.json 2> gpurun_out/r2c5_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2c5_bench.json')); r=d['roofline']
print('value %.3e ms %.3f frac %.3f kernel_ms %.4f e2e %.3e'%(d['value'],d['ms_per_step'],r['frac'],r['kernel_ms'],d['e2e']['value']))"
timeout 300 python scripts/bench_reference_gpu.py --workload c2 > gpurun_out/r2_reference_on_b200_c2.json 2> gpurun_out/r2c5_refgpu.err; cat gpurun_out/r2_reference_on_b200_c2.json
timeout 300 python scripts/bench_reference_gpu.py --workload c3 --steps 3 > gpurun_out/r2_reference_on_b200_c3.json 2>> gpurun_out/r2c5_refgpu.err; cat gpurun_out/r2_reference_on_b200_c3.json
