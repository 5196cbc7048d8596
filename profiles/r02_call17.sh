#!/bin/bash
# Round-2 call 17 (4 GPUs): image-tile scaling with the lazy join on Cornell and BreakfastRoom, config-5 post sweep point
set -u; mkdir -p gpurun_out
t4() { local name=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29546 bench.py --gpus 4 --steps 6 --warmup 3 "$@" 2> gpurun_out/c17_${name}.err | tail -1 > gpurun_out/c17_${name}.json; }
t4 cornell_4gpu --workload cornell_1080p_d8
t4 breakfast_4gpu --workload breakfast_1080p_d8
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c17_*.json")):
    try:
        r = json.loads(open(f).read()); k = r["roofline"]["kernel_ms_per_step"]
        print(f, "N=%d %.1f Mpaths/s  %.2f ms/step  ext %.2f shade %.2f conn %.2f  e2e %.1f clocks %s" % (r["n_gpus"], r["value"], r["ms_per_step"], k["extend"], k["shade"], k["connect"], r["e2e"]["value"], r["clocks"]))
    except Exception as e: print(f, "unreadable", e, open(f.replace(".json", ".err")).read()[-800:])
PY
