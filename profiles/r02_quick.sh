#!/bin/bash
# bench at N vectors (default 250k) + instruction count per expansion of the search kernel (ncu, metrics only)
N=${1:-250000}; shift
mkdir -p gpurun_out
CMD="python bench.py --nvec $N --steps 10 --warmup 3 --no-gate --no-stream --no-recall --no-cpu-baseline $*"
$CMD 2>gpurun_out/err_quick.log | tee gpurun_out/quick_plain.json | python profiles/pj.py && \
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k regex:search_kernel -s 4 -c 1 --csv --log-file gpurun_out/quick_metrics.csv $CMD --inflight 1 --steps 2 --warmup 1 > /dev/null 2>&1
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
if [ -n "$FULL" ]; then
  ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 4 -c 1 -f -o gpurun_out/$FULL $CMD --inflight 1 --steps 2 --warmup 1 > /dev/null 2>&1
  ls -la gpurun_out/$FULL.ncu-rep
fi
