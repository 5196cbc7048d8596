#!/bin/bash
# Round-2 call 12: branch-free triangle test; suite, four workloads, final ncu captures (launch list + full sets of the three Cornell kernels)
set -u; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/c12_pytest.txt
b() { local name=$1; shift; env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --workload $WL 2> gpurun_out/c12_${name}.err | tail -1 > gpurun_out/c12_${name}.json; }
for WL in cornell_1080p_d8 breakfast_1080p_d8 glass_1080sq_d16 viking_1080sq_d8; do b ${WL} B200PT_DEBUG=1; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c12_launches_cornell.csv python bench.py --steps 1 --warmup 3 --frames-per-step 8 --no-cpu-baseline > gpurun_out/c12_ncu_list.log 2>&1
B200PT_OVERLAP=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_shade_hit|k_connect|k_extend" -s 0 -c 6 -o gpurun_out/c12_cornell -f python bench.py --steps 1 --warmup 3 --frames-per-step 8 --no-cpu-baseline > gpurun_out/c12_ncu_full.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c12_*.json")):
    try:
        r = json.loads(open(f).read()); k = r["roofline"]["kernel_ms_per_step"]
        print(f, "%.1f Mpaths/s  %.2f ms/step  ext %.2f shade %.2f conn %.2f  pipe %.3f e2e %.1f" % (r["value"], r["ms_per_step"], k["extend"], k["shade"], k["connect"], r["roofline"]["pipeline"]["frac"], r["e2e"]["value"]))
    except Exception as e: print(f, "unreadable", e)
PY
cat gpurun_out/c12_pytest.txt
