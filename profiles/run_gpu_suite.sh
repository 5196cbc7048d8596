#!/bin/bash
# full GPU parity suite (1 GPU)
cd /root/repo; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|Error:|^E   .*Assert|^FAILED|^E  " | cut -c1-300 | head -40
