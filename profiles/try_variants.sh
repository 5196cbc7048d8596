#!/bin/bash
# usage: try_variants.sh  -> bench each SHADE_MIN_BLOCKS variant (rebuilds on the GPU box)
cd /root/repo
for mb in 3 4 5; do
  (cd vulkan-path-tracer_b200 && touch csrc/wavefront_kernels.cu && make -s -j16 EXTRA=-DSHADE_MIN_BLOCKS=$mb 2>&1 | grep -iE "error" ; grep -A2 "k_shade_hit" build/wavefront_kernels.ptxas.log | grep -oE "[0-9]+ bytes spill stores|Used [0-9]+ registers" | tr '\n' ' ')
  echo " <- SHADE_MIN_BLOCKS=$mb"
  python bench.py --steps 3 --warmup 3 --frames-per-step 32 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('Mpaths/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k:round(v,2) for k,v in d['roofline']['kernel_ms_per_step'].items()})"
done
