#!/bin/bash
N=$1
mkdir -p gpurun_out
for V in "--c4-prefix 65536" "--c4-prefix 16384" "--c4-prefix 32768 --c4-growth 6"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload c4 --steps 5 --warmup 2 --kprime 100 $V 2> gpurun_out/err_c4_$N.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split(chr(10))[-1]); print('$V', round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']))"
done
