#!/bin/bash
# Round-2 evidence run on one B200: GPU tests, the driver's two bench commands, launch list + ncu captures of K3 and K5.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_r02.log
python bench.py --impl reference --gpus 1 --steps 5 --warmup 2 > gpurun_out/bench_r02_ref.json 2> gpurun_out/bench_r02_ref.err
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r02_1gpu.json 2> gpurun_out/bench_r02_1gpu.err
tail -14 gpurun_out/bench_r02_1gpu.err
python profiles/pj.py < gpurun_out/bench_r02_1gpu.json
CMD="python bench.py --steps 2 --warmup 1 --inflight 1 --no-gate --no-stream --no-recall --no-cpu-baseline --no-c4"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 1 --no-gate --no-recall --no-cpu-baseline --c4-n 2000000 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:search_kernel -s 4 -c 1 -f -o gpurun_out/k3_r02_final $CMD > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:exhaustive_scan_tc16 -s 8 -c 1 -f -o gpurun_out/k5_r02_final python bench.py --workload c4 --n 4000000 --steps 1 --warmup 1 --no-recall --kprime 100 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:fastscan_blocks -s 2 -c 1 -f -o gpurun_out/k2_r02_final python bench.py --steps 1 --warmup 1 --no-gate --no-recall --no-cpu-baseline --no-c4 > /dev/null 2>&1
ls -la gpurun_out/*_r02_final.ncu-rep gpurun_out/launches_r02.csv
