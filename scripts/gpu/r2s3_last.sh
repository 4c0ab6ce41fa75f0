#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/r2s3_last_smoke.log
timeout 60 python -u -m pytest tests/test_tc_kernels.py -x -q -m gpu 2>&1 | tail -1
