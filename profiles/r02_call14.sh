#!/bin/bash
# Round-2 call 14 (1 GPU): heterogeneous volumes + atmosphere parity, atmosphere spp sweep, compute-sanitizer on the new paths, whole suite, Cornell bench
set -u; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "heterogeneous or atmosphere" -s 2>&1 | grep -v "^$" | tail -60 > gpurun_out/c14_new.log
timeout 300 python profiles/r02_atm_sweep.py > gpurun_out/r02_atm_sweep.txt 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python profiles/sanitize_small.py > gpurun_out/c14_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/c14_memcheck.log
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -30 > gpurun_out/c14_suite.log
b() { local name=$1; shift; timeout 300 python bench.py --steps 8 --warmup 3 "$@" 2> gpurun_out/c14_${name}.err | tail -1 > gpurun_out/c14_${name}.json; }
b cornell --workload cornell_1080p_d8
b breakfast --workload breakfast_1080p_d8
grep -n "agreement\|rel L2\|passed\|failed\|Error\|assert" gpurun_out/c14_new.log | tail -40; cat gpurun_out/r02_atm_sweep.txt; tail -5 gpurun_out/c14_memcheck.log; tail -8 gpurun_out/c14_suite.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/c14_*.json")):
    try:
        r = json.loads(open(f).read()); k = r["roofline"]["kernel_ms_per_step"]
        print(f, "N=%d %.1f Mpaths/s  %.2f ms/step  ext %.2f shade %.2f conn %.2f  e2e %.1f" % (r["n_gpus"], r["value"], r["ms_per_step"], k["extend"], k["shade"], k["connect"], r["e2e"]["value"]))
    except Exception as e: print(f, "unreadable", e, open(f.replace(".json", ".err")).read()[-800:])
PY
