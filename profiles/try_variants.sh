#!/bin/bash
# usage: bash profiles/try_variants.sh "<EXTRA flags 1>" "<EXTRA flags 2>" ...   (rebuilds on the GPU box, benches each variant)
cd /root/repo
BENCH_ARGS=${BENCH_ARGS:---steps 3 --warmup 3 --frames-per-step 32 --no-cpu-baseline}
for extra in "$@"; do
  (cd vulkan-path-tracer_b200 && touch csrc/*.cu && make -s -j16 EXTRA="$extra" 2>&1 | grep -iE " error" ; grep -A2 "k_shade_hit" build/wavefront_kernels.ptxas.log | grep -oE "[0-9]+ bytes spill stores|Used [0-9]+ registers" | tr '\n' ' ')
  echo " <- EXTRA='$extra'"
  for w in ${WORKLOADS:-cornell_1080p_d8}; do python bench.py $BENCH_ARGS --workload $w 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['workload'], 'Mpaths/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k:round(v,2) for k,v in d['roofline']['kernel_ms_per_step'].items()})"; done
  if [ -n "$PARITY" ]; then python profiles/parity_report.py; fi
done
