#!/bin/bash
# Round-2 call 15 (1 GPU): heterogeneous volumes + walks + atmosphere parity, whole suite, racecheck on the small workloads, bench lines
set -u; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "heterogeneous or atmosphere or walks" -s 2>&1 | grep -v "^$" | tail -80 > gpurun_out/c15_new.log
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -30 > gpurun_out/c15_suite.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python profiles/sanitize_small.py > gpurun_out/c15_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/c15_racecheck.log
b() { local name=$1; shift; timeout 300 python bench.py --steps 8 --warmup 3 "$@" 2> gpurun_out/c15_${name}.err | tail -1 > gpurun_out/c15_${name}.json; }
b cornell --workload cornell_1080p_d8
b breakfast --workload breakfast_1080p_d8
b glass --workload glass_1080sq_d16
b viking --workload viking_1080sq_d8
b post --workload post_4k
grep -n "agreement\|rel L2\|walks case\|passed\|failed\|Error\|assert" gpurun_out/c15_new.log | tail -50; tail -5 gpurun_out/c15_racecheck.log; tail -8 gpurun_out/c15_suite.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c15_*.json")):
    try:
        r = json.loads(open(f).read())
        if "kernel_ms_per_step" in r.get("roofline", {}):
            k = r["roofline"]["kernel_ms_per_step"]
            print(f, "N=%d %.1f Mpaths/s  %.2f ms/step  ext %.2f shade %.2f conn %.2f  e2e %.1f" % (r["n_gpus"], r["value"], r["ms_per_step"], k["extend"], k["shade"], k["connect"], r["e2e"]["value"]))
        else: print(f, r["value"], r["unit"], r["ms_per_step"], r["roofline"]["frac"])
    except Exception as e: print(f, "unreadable", e, open(f.replace(".json", ".err")).read()[-800:])
PY
