#!/bin/bash
# K2 stream A/B on one box: named library variants ("base" = the shipped one) at N vectors, then ncu counters of K2 for the last one.
N=${1:-250000}; shift
mkdir -p gpurun_out
B="python bench.py --nvec $N --steps 5 --warmup 3 --no-gate --no-recall --no-cpu-baseline --no-c4"
for name in "$@"; do
  lib=""; [ "$name" != base ] && lib=$PWD/rabitq-ann-search_b200/cphnsw_b200/variants/libcphnsw_b200_$name.so
  echo "== $name"
  CPHNSW_B200_LIB=$lib $B 2>gpurun_out/err_k2ab.log | tee -a gpurun_out/r02_k2ab.jsonl | python profiles/pj.py || tail -5 gpurun_out/err_k2ab.log
  grep "K2 stream times" gpurun_out/err_k2ab.log
done
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:fastscan_blocks -s 2 -c 1 --csv --log-file gpurun_out/k2ab_metrics.csv $B --steps 2 --warmup 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/k2ab_metrics.csv')) if len(r) > 10]
h = rows[0]; mi, vi = h.index('Metric Name'), h.index('Metric Value')
print({r[mi]: r[vi] for r in rows[1:]})
PY
