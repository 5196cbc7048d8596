#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "post" 2>&1 | grep -E "passed|failed|^E   .*Assert|^FAILED" | cut -c1-300 | head
python bench.py --workload post_4k 2>/dev/null | tail -1 | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/post_launches.csv python bench.py --workload post_4k > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('/root/repo/gpurun_out/post_launches.csv')) if len(r)>5]
h=rows[0]; ki=h.index("Kernel Name"); mi=h.index("Metric Name"); vi=h.index("Metric Value"); ii=h.index("ID")
d={}
for r in rows[1:]:
    d.setdefault((r[ii], r[ki].split('(')[0]), {})[r[mi]]=r[vi]
for (i,k),m in list(d.items())[:24]: print(i,k[:40],m)
PY
