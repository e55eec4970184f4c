#!/bin/bash
# The shipped library at N vectors, then its results against the unmodified reference on the same index (ids as multisets and
# distance bits: the cpu_baseline leg of bench.py).  Named variants (build.py --variant) can be given for an A/B.
N=${1:-250000}; shift
mkdir -p gpurun_out
bash profiles/r02_variants.sh $N base "$@"
python bench.py --nvec $N --steps 5 --warmup 3 --no-gate --no-stream --no-recall --no-c4 --cpu-budget 8 2>gpurun_out/err_coop.log | tee gpurun_out/bench_coop_parity.json | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().split('\n')[-1]); c = d['cpu_baseline']
print('parity vs reference:', c['sample'][:60], '| ids identical', c['ids_identical_to_gpu'], '| distance bits identical', c['distance_bits_identical_to_gpu'], '| qps', round(d['value']))"
