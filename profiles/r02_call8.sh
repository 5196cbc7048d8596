#!/bin/bash
# Round-2 call 8: two wave contexts on two streams (B200PT_OVERLAP), e2e host-time breakdown, full suite
set -u; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/c8_pytest.txt
b() { local name=$1; shift; env "$@" timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --workload $WL 2> gpurun_out/c8_${name}.err | tail -1 > gpurun_out/c8_${name}.json; }
for WL in cornell_1080p_d8 breakfast_1080p_d8 glass_1080sq_d16 viking_1080sq_d8; do b ${WL}_overlap1 B200PT_DEBUG=1; b ${WL}_overlap0 B200PT_OVERLAP=0; done
WL=cornell_1080p_d8; b cornell_e2e_breakdown B200PT_BENCH_E2E_BREAKDOWN=1; grep "e2e" gpurun_out/c8_cornell_e2e_breakdown.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c8_*.json")):
    try:
        r = json.loads(open(f).read()); k = r["roofline"]["kernel_ms_per_step"]
        print(f, "%.1f Mpaths/s  %.2f ms/step  ext %.2f shade %.2f conn %.2f  pipe %.3f e2e %.1f" % (r["value"], r["ms_per_step"], k["extend"], k["shade"], k["connect"], r["roofline"]["pipeline"]["frac"], r["e2e"]["value"]))
    except Exception as e: print(f, "unreadable", e)
PY
cat gpurun_out/c8_pytest.txt
