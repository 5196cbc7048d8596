#!/bin/bash
# round-1 session-3 measurement pass (1 GPU): bench lines for every workload, reference arm, post chain, ncu launch list + full captures
cd /root/repo; mkdir -p gpurun_out
python bench.py > gpurun_out/s3_bench_cornell.json 2> gpurun_out/s3_bench_cornell.err
for w in breakfast_1080p_d8 glass_1080sq_d16 viking_1080sq_d8; do python bench.py --workload $w --steps 4 --warmup 3 --frames-per-step 32 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/s3_bench_$w.json; done
python bench.py --impl reference --steps 3 --warmup 3 2>/dev/null | tail -1 > gpurun_out/s3_bench_reference.json
python bench.py --workload post_4k 2>/dev/null | tail -1 > gpurun_out/s3_bench_post_4k.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/s3_launches_cornell.csv python bench.py --steps 1 --warmup 3 --frames-per-step 8 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/s3_launches_breakfast.csv python bench.py --workload breakfast_1080p_d8 --steps 1 --warmup 3 --frames-per-step 8 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_shade_hit|k_connect|k_extend" -c 6 -o gpurun_out/s3_ncu_cornell -f python bench.py --steps 1 --warmup 3 --frames-per-step 8 --no-cpu-baseline > /dev/null 2>&1
for f in gpurun_out/s3_bench_*.json; do echo "$f: $(cut -c1-230 $f)"; done
