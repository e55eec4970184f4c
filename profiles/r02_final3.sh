#!/bin/bash
# Final round-2 evidence on one B200: the full GPU suite, the driver's two bench commands, then ncu of K3 (1M, one batch at a time).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_r02c.log
S=$(date +%s)
python bench.py --impl reference --gpus 1 --steps 5 --warmup 2 > gpurun_out/bench_r02_ref.json 2> gpurun_out/bench_r02_ref.err
echo "reference arm wall seconds: $(( $(date +%s) - S ))"
S=$(date +%s)
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r02_1gpu.json 2> gpurun_out/bench_r02_1gpu.err
echo "main arm wall seconds: $(( $(date +%s) - S ))"
python profiles/pj.py < gpurun_out/bench_r02_1gpu.json
CMD="python bench.py --steps 2 --warmup 1 --inflight 1 --no-gate --no-stream --no-recall --no-cpu-baseline --no-c4"
ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 4 -c 1 -f -o gpurun_out/k3_r02_final $CMD > /dev/null 2>&1
ls -la gpurun_out/k3_r02_final.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 1 --no-gate --no-recall --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/launches_r02.csv | cut -c1-200
