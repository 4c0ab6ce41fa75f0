#!/bin/bash
N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 200 $TR bench.py --gpus $N --workload c4 --steps 20 --warmup 3 > gpurun_out/r2s3_bench_c4_${N}gpu.out 2> gpurun_out/r2s3_bench_c4_${N}gpu.err
grep "^{" gpurun_out/r2s3_bench_c4_${N}gpu.out | tail -1 > gpurun_out/r2s3_bench_c4_${N}gpu.json
python -c "
import json; d=json.load(open('gpurun_out/r2s3_bench_c4_${N}gpu.json')); print('c4', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['loss'])"
rm -f gpurun_out/r2s3_bench_c4_${N}gpu.out
