#!/bin/bash
N=${1:-250000}
mkdir -p gpurun_out
B="python bench.py --nvec $N --steps 10 --warmup 3 --no-gate --no-stream --no-recall --no-cpu-baseline"
for V in "--inflight 2 --opt ctas_per_sm=12" "--inflight 2 --opt ctas_per_sm=10" "--inflight 2 --opt ctas_per_sm=8" "--inflight 2 --opt ctas_per_sm=6"; do
  echo "== $V"
  $B $V 2>gpurun_out/err_eval.log | tee -a gpurun_out/r02_eval.jsonl | python profiles/pj.py
done
