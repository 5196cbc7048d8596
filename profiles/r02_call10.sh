#!/bin/bash
# Round-2 call 10: lazy join (waves of consecutive path_trace calls overlap), balanced waves, per-context dispatch tables: suite + the four workloads + post
set -u; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/c10_pytest.txt
b() { local name=$1; shift; env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --workload $WL 2> gpurun_out/c10_${name}.err | tail -1 > gpurun_out/c10_${name}.json; }
for WL in cornell_1080p_d8 breakfast_1080p_d8 glass_1080sq_d16 viking_1080sq_d8; do b ${WL} B200PT_DEBUG=1; done
WL=cornell_1080p_d8; b cornell_overlap0 B200PT_OVERLAP=0
timeout 200 python bench.py --workload post_4k 2> gpurun_out/c10_post.err | tail -1 > gpurun_out/c10_post.json
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c10_*.json")):
    try:
        r = json.loads(open(f).read())
        if "kernel_ms_per_step" in r.get("roofline", {}):
            k = r["roofline"]["kernel_ms_per_step"]
            print(f, "%.1f Mpaths/s  %.2f ms/step  ext %.2f shade %.2f conn %.2f  pipe %.3f e2e %.1f wall %.3f" % (r["value"], r["ms_per_step"], k["extend"], k["shade"], k["connect"], r["roofline"]["pipeline"]["frac"], r["e2e"]["value"], r["notes"]["wall_s"]))
        else: print(f, r["value"], r["unit"], r["ms_per_step"], r["roofline"]["frac"])
    except Exception as e: print(f, "unreadable", e)
PY
cat gpurun_out/c10_pytest.txt
