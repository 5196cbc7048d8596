#!/bin/bash
# Round-2 call 19 (1 GPU, final): whole GPU suite + smoke at HEAD; ncu launch list and full sets of the BreakfastRoom traversal kernels (round-2 tree)
set -u; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/c19_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/c19_smoke.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c19_launches_breakfast.csv python bench.py --workload breakfast_1080p_d8 --steps 1 --warmup 3 --frames-per-step 8 --no-cpu-baseline > gpurun_out/c19_ncu_list.log 2>&1
B200PT_OVERLAP=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_extend_dyn|k_shadow_dyn|k_connect|k_shade_hit" -s 0 -c 8 -o gpurun_out/c19_breakfast -f python bench.py --workload breakfast_1080p_d8 --steps 1 --warmup 3 --frames-per-step 8 --no-cpu-baseline > gpurun_out/c19_ncu_full.log 2>&1
cat gpurun_out/c19_suite.log; tail -2 gpurun_out/c19_smoke.log; ls -la gpurun_out/c19_breakfast.ncu-rep; tail -3 gpurun_out/c19_ncu_full.log
