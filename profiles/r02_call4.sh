#!/bin/bash
# Round-2 call 4: ray sort for L2-resident scenes, row-block post pass, config-3 image tests, parity bars at 1e-3.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -s 2>&1 | grep -E "passed|failed|matched-seed|^E |Error" | tail -30 > gpurun_out/c4_pytest.txt
b() { local name=$1; shift; env "$@" timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --workload $WL 2> gpurun_out/c4_${name}.err | tail -1 > gpurun_out/c4_${name}.json; }
for WL in breakfast_1080p_d8 viking_1080sq_d8 glass_1080sq_d16; do b ${WL}_sort1 B200PT_DEBUG=1; b ${WL}_sort0 B200PT_SORT=0; done
WL=cornell_1080p_d8; b cornell_default B200PT_DEBUG=1
timeout 200 python bench.py --workload post_4k 2> gpurun_out/c4_post.err | tail -1 > gpurun_out/c4_post.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 3 2> gpurun_out/c4_reference.err | tail -1 > gpurun_out/c4_reference.json
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c4_*.json")):
    try:
        r = json.loads(open(f).read())
        if "kernel_ms_per_step" in r.get("roofline", {}):
            k = r["roofline"]["kernel_ms_per_step"]
            print(f, "%.1f Mpaths/s  %.2f ms/step  ext %.2f shade %.2f conn %.2f  frac %.3f pipe %.3f e2e %.1f" % (r["value"], r["ms_per_step"], k["extend"], k["shade"], k["connect"], r["roofline"]["frac"], r["roofline"]["pipeline"]["frac"], r["e2e"]["value"]))
        else: print(f, r["value"], r["unit"], r["ms_per_step"], r.get("roofline", {}).get("frac"), r.get("e2e"), r.get("cpu_baseline"))
    except Exception as e: print(f, "unreadable", e)
PY
cat gpurun_out/c4_pytest.txt
