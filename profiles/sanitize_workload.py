#!/usr/bin/env python
"""A small pass over every hot kernel for compute-sanitizer (racecheck / memcheck: one tool per gpurun call):
K1 + K3 (counting build and the fast build, frontier overflow re-run, two batches in flight), K2 (general and streaming
forms), K5 (tcgen05 f16, tcgen05 i8, popcount) and the select / re-rank kernel, the build-side encoder N3.
Results are checked against the oracle so that a sanitizer run is also a parity run."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT / "rabitq-ann-search_b200", ROOT / "oracle", ROOT / "tests"):
    sys.path.insert(0, str(p))
import torch  # noqa: E402

import common  # noqa: E402
from common import co  # noqa: E402
from cphnsw_b200 import hooks  # noqa: E402

oracle = co.Oracle()


def same(ids, d, oid, od):
    gi, gd = common.sorted_rows(ids, d)
    wi, wd = common.sorted_rows(oid, od)
    assert np.array_equal(gi, wi) and np.array_equal(gd.view(np.uint32), wd.view(np.uint32))


for dim, bits in ((96, 4), (128, 1), (256, 2)):
    fab = common.fabricate(1500, dim, bits, seed=dim, layers=1, counts=(32, 31, 9), gamma=1.2, gamma_max=2.0, gamma_warmup=4)
    ix = common.gpu_index_from(fab)
    view = oracle.index_view(fab)
    q = np.random.default_rng(1).standard_normal((48, dim)).astype(np.float32)
    oid, od, _ = oracle.search_batch(view, q, 10)
    same(*ix.search_batch(q, 10), oid, od)                       # counting kernels
    ix.set_option("collect_stats", 0)
    for ids, d in ix.search_batches([q, q], 10):                 # fast kernels, two lanes
        same(ids, d, oid, od)
    ix.set_option("beam_capacity", 64)                           # frontier overflow -> in-stream re-run
    same(*ix.search_batch(q, 10), oid, od)
    prep = hooks.prepare_queries(ix, torch.from_numpy(q[:4]).cuda())
    dqp = torch.full((1500,), 150.0, device="cuda")
    hooks.fastscan_blocks(ix, prep["uplanes"][:1], prep["coeffs"][:1], dqp, first_vertex=0, nblocks=1500, want=("est", "lower"))   # streaming form
    qb = torch.arange(1500, device="cuda", dtype=torch.int32) % 4
    hooks.fastscan_blocks(ix, prep["uplanes"], prep["coeffs"], dqp, first_vertex=0, nblocks=1500, query_of_block=qb)            # general form
    torch.cuda.synchronize()
    print(f"search + fastscan ok: dim {dim}, {bits}-bit")

fab = common.fabricate(3000, 96, 1, seed=5, degenerate=True)
view = oracle.index_view(fab)
q = np.random.default_rng(2).standard_normal((9, 96)).astype(np.float32)
for tc in (2, 1, 0):
    ix = common.gpu_index_from(fab)
    ix.set_option("exhaustive_tensor_cores", tc)
    ids, d = hooks.exhaustive_search(ix, torch.from_numpy(q), 10, 100)
    ids, d = ids.cpu().numpy(), d.cpu().numpy()
    for i in range(len(q)):
        oi, odd, _, _ = oracle.exhaustive(view, fab, q[i], 10, 100)
        assert np.array_equal(ids[i, :len(oi)], oi) and np.array_equal(d[i, :len(oi)].view(np.uint32), np.asarray(odd, np.float32).view(np.uint32))
    print(f"exhaustive scan ok: form {tc}")

import cphnsw_b200  # noqa: E402

vec, pids, nbr = common.neighbor_code_case(96, 40, seed=3)
ids32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.uint32).view(np.int32))  # noqa: E731
for bits in (1, 4):
    for tile in (0, 1):
        ix = cphnsw_b200.CPIndex(96, bits)
        ix.set_option("neighbor_codes_tile", tile)
        want_c, want_a = common.expected_neighbor_codes(oracle, 96, bits, vec, pids, nbr)
        codes, aux = hooks.neighbor_codes(ix, torch.from_numpy(vec), ids32(nbr), ids32(pids))
        assert np.array_equal(codes.cpu().numpy(), want_c) and np.array_equal(aux.cpu().numpy().view(np.uint32), want_a.view(np.uint32))
    print(f"neighbour codes ok: {bits}-bit")
print("sanitize workload done")
