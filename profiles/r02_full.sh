#!/bin/bash
# The driver's own commands on one box: reference arm first (it builds the index), then this arm; then ncu of K3.
mkdir -p gpurun_out
python bench.py --impl reference --gpus 1 --steps 5 --warmup 2 > gpurun_out/bench_r02_ref.json 2> gpurun_out/bench_r02_ref.err
tail -2 gpurun_out/bench_r02_ref.err
python bench.py --gpus 1 --steps 20 --warmup 5 $BENCH_ARGS > gpurun_out/bench_r02_1gpu.json 2> gpurun_out/bench_r02_1gpu.err
tail -12 gpurun_out/bench_r02_1gpu.err
python profiles/pj.py < gpurun_out/bench_r02_1gpu.json
if [ -n "$NCU" ]; then
  CMD="python bench.py --steps 2 --warmup 1 --inflight 1 --no-gate --no-stream --no-recall --no-cpu-baseline"
  ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 4 -c 1 -f -o gpurun_out/$NCU $CMD > /dev/null 2>&1
  ls -la gpurun_out/$NCU.ncu-rep
fi
