#!/bin/bash
# Where does a DB shard's time go?  One GPU scanning what one of 8 shards holds (1.25M of 10M), per-kernel times.
mkdir -p gpurun_out
N=${1:-1250000}
CMD="python bench.py --workload c4 --n $N --steps 3 --warmup 2 --no-recall"
$CMD > gpurun_out/c4_probe_$N.json 2> gpurun_out/c4_probe.err && tail -2 gpurun_out/c4_probe.err && python -c "
import json;d=json.loads(open('gpurun_out/c4_probe_$N.json').read().strip().split('\n')[-1]);print(d['ms_per_step'],d['value'],d['step_ms'],d['roofline']['frac'])" && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c4_launches_$N.csv $CMD > /dev/null 2>&1
python - <<PY
import csv
from collections import OrderedDict
rows=[r for r in csv.reader(open('gpurun_out/c4_launches_$N.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); ui=h.index('Metric Unit')
seq=[(r[ki][:70], float(r[vi].replace(',',''))/ (1 if r[ui]=='us' else 1000 if r[ui]=='ns' else 0.001)) for r in rows[1:]]
# last step = last occurrences
print(len(seq),'launches; tail:')
for n,t in seq[-26:]: print(f'{t:10.1f} us  {n}')
PY
