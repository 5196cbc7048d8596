#!/bin/bash
# Round-2 measurement of the opt-in tree-quality pass (B200PT_BVH_SAH=1): run under gpurun on ONE GPU, e.g.
#   gpurun --timeout 1500 -- 'bash profiles/next_round_sah.sh'      (24 short bench runs + the parity test: ~15 min)
# Model forecast (profiles/bvh_lab.py, profiles/r01_bvh_lab_*.json): BVH4 node visits per ray about -40 % on BreakfastRoom, BVH2 -25..-35 % on viking_room,
# -45 % on CornellBoxGlass; on the 12-triangle Cornell box (headline workload) levels 2 / 3 cut the triangle tests per bounce ray from 7.8 to 2.4.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "sah_rebuild" --runxfail 2>&1 | tail -5 > gpurun_out/sah_parity.txt
for wl in cornell_1080p_d8 glass_1080sq_d16 viking_1080sq_d8 breakfast_1080p_d8; do
  for sah in 0 1 2 3 4 5; do
    B200PT_BVH_SAH=$sah timeout 200 python bench.py --workload $wl --steps 6 --warmup 3 --no-cpu-baseline 2> gpurun_out/sah_${wl}_${sah}.err | tail -1 > gpurun_out/sah_${wl}_${sah}.json
  done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/sah_*_?.json")):
    try: r = json.loads(open(f).read()); print(f, r["value"], r["unit"], r["roofline"]["frac"])
    except Exception as e: print(f, "unreadable", e)
PY
