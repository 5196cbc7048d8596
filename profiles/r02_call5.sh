#!/bin/bash
# does ray coherence help the dynamic-fetch traversal at all?  per-launch durations of k_extend_dyn / k_shadow_dyn with and without the (index) ray sort
set -u; mkdir -p gpurun_out
for s in 0 1; do
  B200PT_SORT=$s timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_extend_dyn|k_shadow_dyn|k_ray_sort|k_shade_hit|k_connect" -c 120 --csv --log-file gpurun_out/c5_launches_breakfast_sort$s.csv python bench.py --workload breakfast_1080p_d8 --steps 1 --warmup 3 --frames-per-step 8 --no-cpu-baseline > gpurun_out/c5_ncu_$s.log 2>&1
done
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "config3 or post_row or fused_bounce or sah_tree" 2>&1 | tail -5
python - <<'PY'
import csv, collections
for s in (0, 1):
    rows = list(csv.reader(l for l in open(f"gpurun_out/c5_launches_breakfast_sort{s}.csv") if l.startswith('"')))
    h = rows[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    per = collections.defaultdict(list)
    for r in rows[1:]:
        per[r[ki].split("(")[0][-40:]].append(float(r[vi].replace(",", "")))
    print("SORT", s)
    for k, v in per.items(): print("   %-42s n=%3d  first 8 launches (us): %s" % (k, len(v), " ".join("%.0f" % (x / 1000 if x > 5000 else x) for x in v[:8])))
PY
