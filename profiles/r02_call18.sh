#!/bin/bash
# Round-2 call 18 (1 GPU): k_connect with next-entry fetch + L2 prefetch of the late-read words (B200PT_CONNECT_PREFETCH build) vs HEAD
set -u; mkdir -p gpurun_out
b() { local name=$1 lib=$2; shift 2; B200PT_LIB=$PWD/vulkan-path-tracer_b200/$lib timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline "$@" 2> gpurun_out/c18_${name}.err | tail -1 > gpurun_out/c18_${name}.json; }
for v in "" _pf; do
  b cornell$v libb200pt$v.so --workload cornell_1080p_d8
  b breakfast$v libb200pt$v.so --workload breakfast_1080p_d8
  b glass$v libb200pt$v.so --workload glass_1080sq_d16
done
B200PT_LIB=$PWD/vulkan-path-tracer_b200/libb200pt_pf.so timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "matched_seed or checkpoint or partition" 2>&1 | tail -3
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c18_*.json")):
    try:
        r = json.loads(open(f).read()); k = r["roofline"]["kernel_ms_per_step"]
        print(f, "%.1f Mpaths/s  %.2f ms/step  ext %.2f shade %.2f conn %.2f" % (r["value"], r["ms_per_step"], k["extend"], k["shade"], k["connect"]))
    except Exception as e: print(f, "unreadable", e, open(f.replace(".json", ".err")).read()[-600:])
PY
