#!/bin/bash
# session 3: 8-GPU lines with the final kernels -- C4 training step (NCCL all-reduce in the log) and C2 inference + sharded pool
N=8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 200 $TR bench.py --gpus $N --workload c4 --steps 20 --warmup 3 > gpurun_out/r2s3_bench_c4_${N}gpu.out 2> gpurun_out/r2s3_bench_c4_${N}gpu.err
grep -v "^{" gpurun_out/r2s3_bench_c4_${N}gpu.out | grep -i -E "AllReduce: opCount 0|nvls|Connected" | head -24 > gpurun_out/r2s3_nccl_c4_${N}gpu.txt
grep "^{" gpurun_out/r2s3_bench_c4_${N}gpu.out | tail -1 > gpurun_out/r2s3_bench_c4_${N}gpu.json
python -c "
import json; d=json.load(open('gpurun_out/r2s3_bench_c4_${N}gpu.json')); print('c4', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['loss'])"
rm -f gpurun_out/r2s3_bench_c4_${N}gpu.out
tail -2 gpurun_out/r2s3_bench_c4_${N}gpu.err | cut -c1-200
timeout 200 $TR bench.py --gpus $N --no-cpu-baseline > gpurun_out/r2s3_bench_c2_${N}gpu.json 2> gpurun_out/r2s3_bench_c2_${N}gpu.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2s3_bench_c2_${N}gpu.json') if l.startswith('{')][-1]); print('c2', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], (d.get('pool') or {}).get('value'))"
tail -2 gpurun_out/r2s3_bench_c2_${N}gpu.err | cut -c1-200
