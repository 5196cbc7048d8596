#!/bin/bash
# Round-2 call 6: full suite with the real reference assets copied to oracle/_ref/assets, wave-size sweep, compute-sanitizer on the small configs
set -u; mkdir -p gpurun_out
ls oracle/_ref/assets | head -3 > gpurun_out/c6_assets.txt
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/c6_pytest.txt
b() { local name=$1; shift; timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --workload $WL "$@" 2> gpurun_out/c6_${name}.err | tail -1 > gpurun_out/c6_${name}.json; }
WL=cornell_1080p_d8; for f in 0 16 32 64; do b cornell_fif$f --frames-in-flight $f; done
WL=breakfast_1080p_d8; for f in 0 16 32; do b breakfast_fif$f --frames-in-flight $f --frames-per-step 64; done
WL=glass_1080sq_d16; for f in 0 32 64; do b glass_fif$f --frames-in-flight $f --frames-per-step 64; done
WL=viking_1080sq_d8; b viking_default
for tool in memcheck racecheck; do timeout 900 compute-sanitizer --tool $tool --print-limit 20 python profiles/sanitize_small.py > gpurun_out/c6_sanitizer_$tool.txt 2>&1; tail -4 gpurun_out/c6_sanitizer_$tool.txt; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c6_*.json")):
    try:
        r = json.loads(open(f).read()); k = r["roofline"]["kernel_ms_per_step"]
        print(f, "%.1f Mpaths/s  %.2f ms/step  ext %.2f shade %.2f conn %.2f  pipe %.3f | %s" % (r["value"], r["ms_per_step"], k["extend"], k["shade"], k["connect"], r["roofline"]["pipeline"]["frac"], r["config"]["env_map"][:20]))
    except Exception as e: print(f, "unreadable", e)
PY
cat gpurun_out/c6_pytest.txt
