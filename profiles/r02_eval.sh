#!/bin/bash
# Round-2 quick evaluation on the GPU box: parity tests, then the c2 bench at N vectors in several launch shapes.
N=${1:-250000}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "not at_scale" 2>&1 | tail -5
B="python bench.py --nvec $N --steps 10 --warmup 3 --no-gate --no-stream --no-recall"
for V in "--inflight 1" "--inflight 2" "--inflight 2 --opt warps_per_cta=4 --opt ctas_per_sm=8" "--inflight 2 --opt warps_per_cta=2 --opt ctas_per_sm=16" "--inflight 2 --opt warps_per_cta=1 --opt ctas_per_sm=32"; do
  echo "== $V"
  $B --no-cpu-baseline $V 2>gpurun_out/err_eval.log | tee -a gpurun_out/r02_eval.jsonl | python profiles/pj.py
done
