#!/bin/bash
# Where does a DB shard's time go?  One GPU scanning what one of W shards holds, in the shard's piece plan; per-kernel times.
mkdir -p gpurun_out
N=${1:-2500000}; PREFIX=${2:-16384}; KP=${3:-100}
CMD="python bench.py --workload c4 --n $N --steps 3 --warmup 2 --no-recall --c4-pieces --c4-prefix $PREFIX --kprime $KP"
$CMD > gpurun_out/c4_probe_$N.json 2> gpurun_out/c4_probe.err && tail -1 gpurun_out/c4_probe.err && python -c "
import json;d=json.loads(open('gpurun_out/c4_probe_$N.json').read().strip().split('\n')[-1]);print(d['ms_per_step'],d['value'],d['step_ms'],d['roofline']['frac'])" && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/c4_launches_$N.csv $CMD > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/c4_launches_$N.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); ui=h.index('Metric Unit')
seq=[(r[ki][:60], float(r[vi].replace(',',''))/ (1 if r[ui]=='us' else 1000 if r[ui]=='ns' else 0.001)) for r in rows[1:]]
# last step: from the last query_prep
last=[i for i,(n,t) in enumerate(seq) if 'query_prep' in n][-1]
tot=0
for n,t in seq[last:]:
    print(f'{t:10.1f} us  {n}'); tot+=t
print('sum', tot)
PY
