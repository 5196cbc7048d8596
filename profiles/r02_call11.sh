#!/bin/bash
# Round-2 call 11 (2 GPUs): image-tile bench with asynchronous path_trace (gather of step k beside the first bounces of step k+1)
set -u; mkdir -p gpurun_out
t2() { local name=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 6 --warmup 3 "$@" 2> gpurun_out/c11_${name}.err | tail -1 > gpurun_out/c11_${name}.json; }
t2 cornell_2gpu --workload cornell_1080p_d8
t2 breakfast_2gpu --workload breakfast_1080p_d8
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c11_*.json")):
    try:
        r = json.loads(open(f).read()); k = r["roofline"]["kernel_ms_per_step"]
        print(f, "N=%d %.1f Mpaths/s  %.2f ms/step  e2e %.1f wall %.3f clocks %s" % (r["n_gpus"], r["value"], r["ms_per_step"], r["e2e"]["value"], r["notes"]["wall_s"], r["clocks"]))
    except Exception as e: print(f, "unreadable", e, open(f.replace(".json", ".err")).read()[-800:])
PY
