#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./scripts/micro/group_phase_bench 2>&1 | tee gpurun_out/r2c31_group_phase.log
