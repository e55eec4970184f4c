#!/bin/bash
# Quick evaluation on the GPU box: parity tests, bench at N (default 250000), instruction count of the search kernel.
N=${1:-250000}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
CMD="python bench.py --nvec $N --steps 3 --warmup 1 --no-cpu-baseline --no-recall"
$CMD 2>/dev/null | tee gpurun_out/quick_plain.json | python profiles/pj.py && \
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k regex:search_kernel -s 2 -c 1 --csv --log-file gpurun_out/quick_metrics.csv $CMD > /dev/null 2>&1
python - <<'PY'
import csv, json
rows = [r for r in csv.reader(open('gpurun_out/quick_metrics.csv')) if len(r) > 10]
h = rows[0]; mi, vi = h.index('Metric Name'), h.index('Metric Value')
m = {r[mi]: float(r[vi].replace(',', '')) for r in rows[1:]}
d = json.loads(open('gpurun_out/quick_plain.json').read().strip().split('\n')[-1])
nexp = d['roofline']['expansions_per_query'] * 10000
print({k: round(v, 2) for k, v in m.items()})
print('inst/expansion', round(m['smsp__inst_executed.sum'] / nexp, 1), 'dram bytes/expansion', round((m['dram__bytes_read.sum'] + m['dram__bytes_write.sum']) / nexp))
PY
