import sys, time
sys.path.insert(0, 'rabitq-ann-search_b200'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'oracle')
import numpy as np, torch
import common
from cphnsw_b200 import hooks
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
fab = common.fabricate(n, 128, 1, seed=1)
ix = common.gpu_index_from(fab)
q = torch.randn(nq, 128, device='cuda')
res = {}
for tc in (1, 0):
    ix.set_option("exhaustive_tensor_cores", tc)
    for it in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        ids, d = hooks.exhaustive_search(ix, q, 10, 100)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    res[tc] = (ids.cpu().numpy(), d.cpu().numpy())
    print('tensor_cores', tc, 'ms', round(ms, 2), 'qps', round(nq / ms * 1e3), 'pairs/s', f"{n * nq / ms * 1e3:.3e}")
print('identical', np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1].view(np.uint32), res[1][1].view(np.uint32)))
