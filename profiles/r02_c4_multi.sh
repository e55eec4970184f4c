#!/bin/bash
# c4 (10M x 96, DB-sharded) on the box's GPUs: N = 1 and N = $1, k' = 100 and 256
N=$1
mkdir -p gpurun_out
for KP in 100 256; do
  python bench.py --workload c4 --steps 5 --warmup 2 --kprime $KP $( [ $KP = 100 ] && echo --no-recall ) > gpurun_out/bench_c4_r02_kp${KP}_1gpu.json 2> gpurun_out/err_c4.log
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload c4 --steps 5 --warmup 2 --kprime $KP > gpurun_out/bench_c4_r02_kp${KP}_${N}gpu.json 2> gpurun_out/err_c4_$N.log
  tail -1 gpurun_out/err_c4_$N.log
  python - <<PY
import json
for n in (1, $N):
    d = json.loads(open(f"gpurun_out/bench_c4_r02_kp${KP}_{n}gpu.json").read().strip().split("\n")[-1])
    print("k'=$KP N=%d: %d QPS, %.2f ms/step, e2e %d, frac %.3f" % (n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"]))
PY
done
