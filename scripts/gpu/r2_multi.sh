#!/bin/bash
# usage: r2_multi.sh N  -- multi-GPU lines of round 2 (one node, torchrun): c2 (+ sharded pool), c4 training, c5 sweep
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
last() { grep "^{" "$1" | tail -1; }
timeout 240 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_c2_${N}gpu.out 2> gpurun_out/r2_bench_c2_${N}gpu.err
last gpurun_out/r2_bench_c2_${N}gpu.out > gpurun_out/r2_bench_c2_${N}gpu.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c2_${N}gpu.json')); print('c2', d['n_gpus'], d['value'], d['ms_per_step']); p=d['pool']; print('pool', p['value'], p['sweep_ms'], p['per_rank']['worst_over_mean_busy'], p['per_rank']['busy_ms'])"
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 240 $TR bench.py --gpus $N --workload c4 --steps 20 --warmup 3 > gpurun_out/r2_bench_c4_${N}gpu.out 2> gpurun_out/r2_bench_c4_${N}gpu.err
grep -v "^{" gpurun_out/r2_bench_c4_${N}gpu.out | grep -i -E "AllReduce: opCount 0|nvls|Connected" | head -24 > gpurun_out/r2_nccl_c4_${N}gpu.txt
last gpurun_out/r2_bench_c4_${N}gpu.out > gpurun_out/r2_bench_c4_${N}gpu.json
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c4_${N}gpu.json')); print('c4', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['loss'])"
timeout 400 $TR scripts/design_sweep.py --check > gpurun_out/r2_design_sweep_${N}gpu.out 2> gpurun_out/r2_design_sweep_${N}gpu.err
last gpurun_out/r2_design_sweep_${N}gpu.out > gpurun_out/r2_design_sweep_${N}gpu.json
python -c "
import json; d=json.load(open('gpurun_out/r2_design_sweep_${N}gpu.json')); print('c5', d['n_gpus'], d['value'], d['sweep_ms'], d['argmax_agreement_with_oracle_T0'], d['per_rank']['worst_over_mean_busy'])"
rm -f gpurun_out/r2_bench_c4_${N}gpu.out
