#!/bin/bash
# Round-2 call 2: parity suite on the class-queue / fused-bounce build, then A/B of the pipeline shapes and register budgets, then ncu.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/c2_pytest.txt
P=vulkan-path-tracer_b200
b() { # name, env..., -- args
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline ${WL:+--workload $WL} 2> gpurun_out/c2_${name}.err | tail -1 > gpurun_out/c2_${name}.json
}
WL=cornell_1080p_d8
b cornell_fuse2 B200PT_DEBUG=1
b cornell_fuse1 B200PT_FUSE=1
b cornell_fuse0 B200PT_FUSE=0
b cornell_fuse0_noclass B200PT_FUSE=0 B200PT_CLASSES=0
b cornell_fuse2_noclass B200PT_CLASSES=0
b cornell_fuse2_b4 B200PT_LIB=$PWD/$P/libb200pt_b4.so
b cornell_fuse2_b3 B200PT_LIB=$PWD/$P/libb200pt_b3.so
for WL in glass_1080sq_d16 viking_1080sq_d8 breakfast_1080p_d8; do b ${WL}_default B200PT_DEBUG=1; b ${WL}_noclass B200PT_CLASSES=0; done
WL=cornell_1080p_d8
# ncu: launch list + one full capture of the fused bounce kernel (bounce 1 of a wave: skip the bounce-0 launch)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/c2_launches_cornell.csv python bench.py --steps 1 --warmup 3 --frames-per-step 8 --no-cpu-baseline > gpurun_out/c2_ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade_hit -s 0 -c 2 -o gpurun_out/c2_bounce -f python bench.py --steps 1 --warmup 3 --frames-per-step 8 --no-cpu-baseline > gpurun_out/c2_ncu_full.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c2_*.json")):
    try:
        r = json.loads(open(f).read()); k = r["roofline"]["kernel_ms_per_step"]
        print(f, "%.1f Mpaths/s  %.2f ms/step  ext %.2f shade %.2f conn %.2f  frac %.3f pipe %.3f" % (r["value"], r["ms_per_step"], k["extend"], k["shade"], k["connect"], r["roofline"]["frac"], r["roofline"]["pipeline"]["frac"]))
    except Exception as e: print(f, "unreadable", e)
PY
cat gpurun_out/c2_pytest.txt
