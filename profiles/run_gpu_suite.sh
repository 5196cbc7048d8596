#!/bin/bash
# full GPU parity suite + a short bench of every workload (1 GPU)
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|Error:|^E   .*Assert|^FAILED" | cut -c1-300 | head -30
bash profiles/sweep_trav.sh cornell_1080p_d8:B200PT_X=0 breakfast_1080p_d8:B200PT_X=0 2>&1 | grep -v "^$"
