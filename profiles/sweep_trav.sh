#!/bin/bash
# usage (on the GPU box): bash profiles/sweep_trav.sh "workload:ENV=1,ENV2=x" ...   -- A/B of the traversal shapes through env switches (no rebuild)
cd /root/repo
BENCH_ARGS=${BENCH_ARGS:---steps 3 --warmup 3 --frames-per-step 32 --no-cpu-baseline}
run() {  # $1 workload, rest: env assignments
  w=$1; shift
  line=$(env "$@" timeout 300 python bench.py $BENCH_ARGS --workload $w 2>&1 | tail -1)
  echo "$line" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w', '$*', 'Mpaths/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), {k:round(v,2) for k,v in d['roofline']['kernel_ms_per_step'].items()})" 2>/dev/null || echo "$w $* FAILED: $line" | cut -c1-600
}
for spec in "$@"; do
  w=${spec%%:*}; envs=${spec#*:}
  run $w ${envs//,/ }
done
