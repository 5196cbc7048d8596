#!/bin/bash
# full GPU parity suite + post-chain bench (fused / pass-per-pass) + a short bench of the headline workload (1 GPU)
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|Error:|^E   .*Assert|^FAILED" | cut -c1-300 | head -30
python bench.py --workload post_4k 2>/dev/null | tail -1 | cut -c1-400
B200PT_POST_FUSED=0 python bench.py --workload post_4k 2>/dev/null | tail -1 | cut -c1-400
bash profiles/sweep_trav.sh cornell_1080p_d8:B200PT_X=0 2>&1 | grep -v "^$"
