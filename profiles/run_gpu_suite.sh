#!/bin/bash
# full GPU parity suite + node-format sweep + post-chain bench (1 GPU)
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|Error:|^E   .*Assert|^FAILED" | cut -c1-300 | head -30
bash profiles/sweep_trav.sh breakfast_1080p_d8:B200PT_NODES=half breakfast_1080p_d8:B200PT_NODES=bvh4 breakfast_1080p_d8:B200PT_NODES=bvh2 viking_1080sq_d8:B200PT_NODES=half viking_1080sq_d8:B200PT_NODES=bvh2 glass_1080sq_d16:B200PT_NODES=half glass_1080sq_d16:B200PT_NODES=bvh2 2>&1 | grep -v "^$"
python bench.py --workload post_4k 2>/dev/null | tail -1 | cut -c1-260
