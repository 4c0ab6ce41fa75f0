#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanitize_smoke.py > gpurun_out/r2c29_sanitize.log 2>&1; echo "sanitizer rc $?"; grep -E "ERROR SUMMARY|rdesign|rnampnn|collate|forest|attention|Invalid|error" gpurun_out/r2c29_sanitize.log | head -30
