#!/bin/bash
# Round-2 call 3: flat (no-hierarchy) traversal of tiny scenes, uniform-class k_extend, big-smem mode (glass), BVH4 treelet (BreakfastRoom), cluster post kernel.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/c3_pytest.txt
b() { local name=$1; shift; env "$@" timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --workload $WL 2> gpurun_out/c3_${name}.err | tail -1 > gpurun_out/c3_${name}.json; }
WL=cornell_1080p_d8
b cornell_default B200PT_DEBUG=1
b cornell_noflat B200PT_FLAT_MAX=0
b cornell_fuse2 B200PT_FUSE=2
b cornell_fuse2_noflat B200PT_FUSE=2 B200PT_FLAT_MAX=0
WL=glass_1080sq_d16
b glass_default B200PT_DEBUG=1
b glass_nobig B200PT_SMEM_BIG=0
b glass_big_fuse2 B200PT_FUSE=2
WL=viking_1080sq_d8
b viking_default B200PT_DEBUG=1
WL=breakfast_1080p_d8
b breakfast_top0 B200PT_TOP_KB=0
b breakfast_top40 B200PT_DEBUG=1
b breakfast_top80 B200PT_TOP_KB=80
b breakfast_top120 B200PT_TOP_KB=120
for v in default:B200PT_DEBUG=1 nosmall:B200PT_POST_SMALL=0; do env ${v#*:} timeout 200 python bench.py --workload post_4k 2> gpurun_out/c3_post_${v%%:*}.err | tail -1 > gpurun_out/c3_post_${v%%:*}.json; done
timeout 900 python profiles/r02_parity_sweep.py > gpurun_out/c3_parity_sweep.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/c3_launches_cornell.csv python bench.py --steps 1 --warmup 3 --frames-per-step 8 --no-cpu-baseline > gpurun_out/c3_ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_shade_hit|k_connect|k_extend" -s 0 -c 6 -o gpurun_out/c3_cornell -f python bench.py --steps 1 --warmup 3 --frames-per-step 8 --no-cpu-baseline > gpurun_out/c3_ncu_full.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c3_*.json")):
    try:
        r = json.loads(open(f).read())
        if "kernel_ms_per_step" in r.get("roofline", {}):
            k = r["roofline"]["kernel_ms_per_step"]
            print(f, "%.1f Mpaths/s  %.2f ms/step  ext %.2f shade %.2f conn %.2f  frac %.3f pipe %.3f" % (r["value"], r["ms_per_step"], k["extend"], k["shade"], k["connect"], r["roofline"]["frac"], r["roofline"]["pipeline"]["frac"]))
        else: print(f, r["value"], r["unit"], r["ms_per_step"], r["roofline"]["frac"])
    except Exception as e: print(f, "unreadable", e)
PY
cat gpurun_out/c3_pytest.txt; tail -60 gpurun_out/c3_parity_sweep.txt
