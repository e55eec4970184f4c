"""K5 at BASELINE config 4's shape: n x 96 (padded to 128), 1-bit per-vertex RaBitQ codes, exhaustive batched scan.

    python profiles/k5_bench.py [n=10000000] [nq=10000] [kprime=100] [--popcount] [--clusters C]

The reference cannot build a 10M index in any reasonable time (and has no exhaustive mode), and the
scan reads only the per-vertex codes, so this script makes the index itself: vectors are encoded exactly as
RaBitQEncoder::encode_impl does (encoder/rabitq_encoder.hpp:225-262: centre, rotate with the index's own
3-layer Hadamard rotation -- K1 run on the base vectors --, sign bits, nop = |x - c|, ip_qo = |rot|_1 / sqrt(D)),
neighbour blocks are left empty (count = 0; the scan never touches them).  Reports time per batch, pairs/s,
int8 tensor throughput and recall@10 against exact brute force on a sample of the queries.
"""
import sys
import time

sys.path.insert(0, 'rabitq-ann-search_b200'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'oracle')
import numpy as np
import torch

import common
from common import co
from cphnsw_b200 import hooks
import cphnsw_b200

args = [a for i, a in enumerate(sys.argv[1:], 1) if not a.startswith('--') and sys.argv[i - 1] not in ('--modes', '--clusters')]
n = int(args[0]) if len(args) > 0 else 10_000_000
nq = int(args[1]) if len(args) > 1 else 10_000
kp = int(args[2]) if len(args) > 2 else 100
clusters = int(sys.argv[sys.argv.index('--clusters') + 1]) if '--clusters' in sys.argv else 0
dim, D, k = 96, 128, 10
dev = torch.device('cuda')
g = torch.Generator(device=dev); g.manual_seed(1234)

t0 = time.time()
centers = torch.randn((max(clusters, 1), dim), generator=g, device=dev) * 4.0
def gen(m):
    x = torch.randn((m, dim), generator=g, device=dev)
    if clusters:
        x += centers[torch.randint(0, clusters, (m,), generator=g, device=dev)]
    return x
base = torch.empty((n, dim), dtype=torch.float32, device=dev)
for s in range(0, n, 1 << 20):
    base[s:s + (1 << 20)] = gen(min(1 << 20, n - s))
centroid = base.mean(0)

# a 64-vertex dummy index with this centroid: gives us K1 (the index's rotation) for the base vectors
tiny = common.fabricate(64, dim, 1, seed=1)
tiny.centroid = centroid.cpu().numpy()
ixt = common.gpu_index_from(tiny)
lay = co.nb_layout(D, 1)
nb_off = co.code_bytes(D, 1)
rec = nb_off + lay["size"]
storage = -(-(8 * ((D + 63) // 64)) // 64) * 64
sd = np.zeros((n, rec), np.uint8)
raw = np.zeros((n, D), np.float32)
norm_sq = np.empty(n, np.float32)
ids_off = nb_off + lay["ids"]
step = 1 << 18
for s in range(0, n, step):
    x = base[s:s + step]
    rot = hooks.prepare_queries(ixt, x, center=True)["rotated"]      # (x - c) rotated and scaled by 1/(D sqrt D)
    nop = (x - centroid).norm(dim=1)
    bits = (rot >= 0).to(torch.uint8).reshape(-1, D // 8, 8)
    packed = (bits << torch.arange(8, device=dev, dtype=torch.uint8)).sum(2).to(torch.uint8)   # bit i of byte j = dim 8j+i
    ipqo = rot.abs().sum(1) / (nop.clamp_min(1e-30) * (D ** 0.5))
    e = s + x.shape[0]
    sd[s:e, :D // 8] = packed.cpu().numpy()
    sd[s:e, storage:storage + 4] = nop.cpu().numpy().view(np.uint8).reshape(-1, 4)
    sd[s:e, storage + 4:storage + 8] = ipqo.to(torch.float32).cpu().numpy().view(np.uint8).reshape(-1, 4)
    sd[s:e, ids_off:ids_off + 128] = 0xFF
    raw[s:e, :dim] = x.cpu().numpy()
    norm_sq[s:e] = (x * x).sum(1).cpu().numpy()
del ixt
print(f"[k5_bench] encoded {n} x {dim} in {time.time() - t0:.1f} s (host arrays {sd.nbytes / 2**30:.1f} + {raw.nbytes / 2**30:.1f} GiB)", flush=True)

t0 = time.time()
ix = cphnsw_b200.CPIndex(dim, 1)
hooks.upload_arrays(ix, D=D, bits=1, dim=dim, search_data=sd, raw=raw, norm_sq=norm_sq, calibration=tiny.calibration,
                    centroid=tiny.centroid, max_level=0, entry_point=0, graph_entry_point=0, rotation_seed=42, layers=[])
del sd, raw
print(f"[k5_bench] index on the device in {time.time() - t0:.1f} s, {ix.info()['device_bytes'] / 2**30:.1f} GiB", flush=True)

q = gen(nq)
res = {}
modes = (1, 0) if '--popcount' in sys.argv else (1,)
if '--modes' in sys.argv:
    modes = tuple(int(x) for x in sys.argv[sys.argv.index('--modes') + 1].split(','))
for tc in modes:
    ix.set_option("exhaustive_tensor_cores", tc)
    times = []
    for it in range(4):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        ids, d = hooks.exhaustive_search(ix, q, k, kp)
        e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = float(np.mean(times[1:]))
    res[tc] = (ids.cpu().numpy(), d.cpu().numpy())
    pairs = n * nq / ms * 1e3
    print(f"[k5_bench] tensor_cores={tc} n={n} nq={nq} k'={kp}: {ms:.2f} ms/batch ({[round(t, 2) for t in times]}), {nq / ms * 1e3:.0f} QPS, "
          f"{pairs:.3e} pairs/s, {pairs * 2 * D / 1e12:.1f} int8 TOP/s", flush=True)
if len(res) >= 2:
    ks = list(res)
    print("[k5_bench] identical:", all(np.array_equal(res[ks[0]][0], res[k_][0]) and np.array_equal(res[ks[0]][1].view(np.uint32), res[k_][1].view(np.uint32)) for k_ in ks[1:]))

# recall@10 on a sample: exact brute force
ns = min(nq, 500)
qs = q[:ns]
best_d = torch.full((ns, k), float('inf'), device=dev)
best_i = torch.zeros((ns, k), dtype=torch.int64, device=dev)
for s in range(0, n, 1 << 20):
    b = base[s:s + (1 << 20)]
    dd = (qs * qs).sum(1, keepdim=True) - 2.0 * (qs @ b.T) + (b * b).sum(1)[None, :]
    td, ti = torch.topk(dd, k, dim=1, largest=False)
    cd, ci = torch.cat([best_d, td], 1), torch.cat([best_i, ti + s], 1)
    best_d, sel = torch.topk(cd, k, dim=1, largest=False)
    best_i = torch.gather(ci, 1, sel)
gt = best_i.cpu().numpy()
got = res[modes[0]][0][:ns]
print(f"[k5_bench] recall@10 (k'={kp}, {ns} queries): {np.mean([len(set(a.tolist()) & set(b.tolist())) / k for a, b in zip(got, gt)]):.4f}")
