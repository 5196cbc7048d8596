#!/bin/bash
# Round-2 call 9 (8 GPUs): image-tile scaling on Cornell and BreakfastRoom, config-5 post sweep
set -u; mkdir -p gpurun_out
t8() { local name=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 6 --warmup 3 "$@" 2> gpurun_out/c9_${name}.err | tail -1 > gpurun_out/c9_${name}.json; }
t8 cornell_8gpu --workload cornell_1080p_d8
t8 breakfast_8gpu --workload breakfast_1080p_d8
t8 post_8gpu --workload post_4k
t4() { local name=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 4 --steps 6 --warmup 3 "$@" 2> gpurun_out/c9_${name}.err | tail -1 > gpurun_out/c9_${name}.json; }
t4 post_4gpu --workload post_4k
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c9_*.json")):
    try:
        r = json.loads(open(f).read())
        if "kernel_ms_per_step" in r.get("roofline", {}):
            k = r["roofline"]["kernel_ms_per_step"]
            print(f, "N=%d %.1f Mpaths/s  %.2f ms/step  ext %.2f shade %.2f conn %.2f  e2e %.1f clocks %s" % (r["n_gpus"], r["value"], r["ms_per_step"], k["extend"], k["shade"], k["connect"], r["e2e"]["value"], r["clocks"]))
        else: print(f, "N=%d" % r["n_gpus"], r["value"], r["unit"], r["ms_per_step"], r["roofline"]["frac"], r["e2e"]["value"])
    except Exception as e: print(f, "unreadable", e, open(f.replace(".json", ".err")).read()[-800:])
PY
