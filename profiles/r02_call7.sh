#!/bin/bash
# Round-2 call 7 (2 GPUs): auto wave size on 1 GPU, then the 2-GPU paths: image-tile bench, config-5 post sweep
set -u; mkdir -p gpurun_out
b() { local name=$1; shift; timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline "$@" 2> gpurun_out/c7_${name}.err | tail -1 > gpurun_out/c7_${name}.json; }
b cornell_auto --workload cornell_1080p_d8
b breakfast_auto --workload breakfast_1080p_d8
b breakfast_fif32 --workload breakfast_1080p_d8 --frames-in-flight 32
b glass_auto --workload glass_1080sq_d16
b viking_auto --workload viking_1080sq_d8
timeout 200 python bench.py --workload post_4k 2> gpurun_out/c7_post_1gpu.err | tail -1 > gpurun_out/c7_post_1gpu.json
t2() { local name=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 4 --warmup 3 "$@" 2> gpurun_out/c7_${name}.err | tail -1 > gpurun_out/c7_${name}.json; }
t2 cornell_2gpu --workload cornell_1080p_d8
t2 breakfast_2gpu --workload breakfast_1080p_d8
t2 post_2gpu --workload post_4k
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c7_*.json")):
    try:
        r = json.loads(open(f).read())
        if "kernel_ms_per_step" in r.get("roofline", {}):
            k = r["roofline"]["kernel_ms_per_step"]
            print(f, "N=%d %.1f Mpaths/s  %.2f ms/step  ext %.2f shade %.2f conn %.2f  pipe %.3f e2e %.1f clocks %s" % (r["n_gpus"], r["value"], r["ms_per_step"], k["extend"], k["shade"], k["connect"], r["roofline"]["pipeline"]["frac"], r["e2e"]["value"], r["clocks"].get("sm_mhz")))
        else: print(f, "N=%d" % r["n_gpus"], r["value"], r["unit"], r["ms_per_step"], r["roofline"]["frac"], r["e2e"]["value"])
    except Exception as e: print(f, "unreadable", e, open(f.replace(".json", ".err")).read()[-600:])
PY
