#!/bin/bash
# round 2, GPU call 1: corrected TMEM microbenchmark, full GPU test suite (incl. the 234-structure valid set),
# baseline bench line, edge-kernel timeline, ncu --set full with source of the shipped edge kernel
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2c1_smi.txt
./scripts/micro/tmem_bench2 > gpurun_out/r2_tmem_bench2.log 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/r2c1_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2c1_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err
python scripts/edge_timeline.py > gpurun_out/r2c1_edge_tl.log 2>&1
python scripts/edge_launch.py 5 > gpurun_out/r2c1_edge_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:edge_mlp_tc -s 2 -c 1 -f -o gpurun_out/r2c1_edge_base \
    python scripts/edge_launch.py 5 > gpurun_out/r2c1_ncu.log 2>&1
tail -3 gpurun_out/r2c1_pytest.log; cat gpurun_out/r2_tmem_bench2.log; cat gpurun_out/r2c1_bench.json; cat gpurun_out/r2c1_edge_plain.log
