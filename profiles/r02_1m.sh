#!/bin/bash
# c2 at 1M with a given set of bench args (index built once per box)
mkdir -p gpurun_out
for V in "$@"; do
  echo "== $V"
  python bench.py --steps 10 --warmup 3 --no-gate --no-stream --no-recall --no-cpu-baseline $V 2>gpurun_out/err_1m.log | tee -a gpurun_out/r02_1m.jsonl | python profiles/pj.py
done
