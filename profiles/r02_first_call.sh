#!/bin/bash
# Round-2 first GPU call: (1) vulkaninfo / host probe, (2) the opt-in SAH tree passes: parity + bench at levels 0..5 on the four workloads.
set -u
mkdir -p gpurun_out
{ echo "== vulkaninfo"; (which vulkaninfo && vulkaninfo --summary) 2>&1 | head -20; echo "== libvulkan"; ldconfig -p | grep -i vulkan; ls /usr/share/vulkan/icd.d /etc/vulkan/icd.d 2>&1;
  echo "== nvidia-smi"; nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv; echo "== host"; nproc; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket"; cat /sys/fs/cgroup/cpu.max 2>&1; } > gpurun_out/r02_probe.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "sah_rebuild" --runxfail 2>&1 | tail -5 > gpurun_out/sah_parity.txt
for wl in cornell_1080p_d8 glass_1080sq_d16 viking_1080sq_d8 breakfast_1080p_d8; do
  for sah in 0 1 2 3 4 5; do
    B200PT_DEBUG=1 B200PT_BVH_SAH=$sah timeout 200 python bench.py --workload $wl --steps 4 --warmup 3 --no-cpu-baseline 2> gpurun_out/sah_${wl}_${sah}.err | tail -1 > gpurun_out/sah_${wl}_${sah}.json
  done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/sah_*_?.json")):
    try: r = json.loads(open(f).read()); print(f, r["value"], r["unit"], r["roofline"]["frac"])
    except Exception as e: print(f, "unreadable", e)
PY
