#!/bin/bash
# Round-2 call 13 (1 GPU): atmosphere parity first, then the whole GPU suite, then the four bench workloads (tri_test<EARLY>)
set -u; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "atmosphere" -s 2>&1 | tail -40 > gpurun_out/c13_atm.log
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -30 > gpurun_out/c13_suite.log
b() { local name=$1; shift; timeout 300 python bench.py --steps 8 --warmup 3 "$@" 2> gpurun_out/c13_${name}.err | tail -1 > gpurun_out/c13_${name}.json; }
b cornell --workload cornell_1080p_d8
b breakfast --workload breakfast_1080p_d8
b glass --workload glass_1080sq_d16
b viking --workload viking_1080sq_d8
b post --workload post_4k
cat gpurun_out/c13_atm.log; cat gpurun_out/c13_suite.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c13_*.json")):
    try:
        r = json.loads(open(f).read())
        if "kernel_ms_per_step" in r.get("roofline", {}):
            k = r["roofline"]["kernel_ms_per_step"]
            print(f, "N=%d %.1f Mpaths/s  %.2f ms/step  ext %.2f shade %.2f conn %.2f  e2e %.1f clocks %s" % (r["n_gpus"], r["value"], r["ms_per_step"], k["extend"], k["shade"], k["connect"], r["e2e"]["value"], r["clocks"]))
        else: print(f, "N=%d" % r["n_gpus"], r["value"], r["unit"], r["ms_per_step"], r["roofline"]["frac"], r["e2e"]["value"])
    except Exception as e: print(f, "unreadable", e, open(f.replace(".json", ".err")).read()[-800:])
PY
