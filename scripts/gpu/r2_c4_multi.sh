#!/bin/bash
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 240 $TR bench.py --gpus $N --workload c4 --steps 20 --warmup 3 > gpurun_out/r2_bench_c4_${N}gpu.json 2> gpurun_out/r2_bench_c4_${N}gpu.err
grep -v "^{" gpurun_out/r2_bench_c4_${N}gpu.json | grep -i -E "AllReduce|nvls|algo|Connected" | head -20 > gpurun_out/r2_nccl_c4_${N}gpu.txt
grep "^{" gpurun_out/r2_bench_c4_${N}gpu.json > gpurun_out/r2_bench_c4_${N}gpu.line.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c4_${N}gpu.line.json')); print('c4', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['loss'])"
tail -3 gpurun_out/r2_bench_c4_${N}gpu.err | cut -c1-200; head -8 gpurun_out/r2_nccl_c4_${N}gpu.txt | cut -c1-200
