#!/bin/bash
# Round-2 call 20 (1 GPU): compute-sanitizer initcheck + synccheck on the small workloads (every kernel family, atmosphere and grid volumes included)
set -u; mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool initcheck --error-exitcode 9 python profiles/sanitize_small.py > gpurun_out/c20_initcheck.log 2>&1; echo "initcheck rc=$?" >> gpurun_out/c20_initcheck.log
timeout 400 compute-sanitizer --tool synccheck --error-exitcode 9 python profiles/sanitize_small.py > gpurun_out/c20_synccheck.log 2>&1; echo "synccheck rc=$?" >> gpurun_out/c20_synccheck.log
tail -4 gpurun_out/c20_initcheck.log; tail -4 gpurun_out/c20_synccheck.log; grep -c "Uninitialized" gpurun_out/c20_initcheck.log
