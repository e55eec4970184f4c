import sys, time
sys.path.insert(0, 'rabitq-ann-search_b200'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'oracle')
import numpy as np, torch
import common
from cphnsw_b200 import hooks
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
fab = common.fabricate(n, 128, 4, seed=1)
ix = common.gpu_index_from(fab)
q = torch.randn(1, 128, device='cuda')
prep = hooks.prepare_queries(ix, q)
dqp = torch.full((n,), 200.0, device='cuda')
for want in (("est", "lower"), ()):
    for it in range(4):
        torch.cuda.synchronize(); t = time.perf_counter()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        out = hooks.fastscan_blocks(ix, prep["uplanes"], prep["coeffs"], dqp, first_vertex=0, nblocks=n, want=want)
        e1.record(); torch.cuda.synchronize()
        print(want, it, 'event ms', e0.elapsed_time(e1), 'wall ms', (time.perf_counter() - t) * 1e3, 'GB/s', n * 2560 / e0.elapsed_time(e1) / 1e6)
