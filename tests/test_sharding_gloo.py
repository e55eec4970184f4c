"""CPU, world_size 2 over gloo: the N>1 plumbing (query sharding, DB sharding + top-k merge) with the
oracle standing in for the per-rank device search."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

import common
from common import co


def _worker(rank, world, port, mode, out):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, str(common.ROOT / "rabitq-ann-search_b200"))
    from cphnsw_b200 import sharding

    o = co.Oracle()
    fab = common.fabricate(600, 32, 1, seed=3, layers=1)
    view = o.index_view(fab)
    q = np.random.default_rng(2).standard_normal((21, 32)).astype(np.float32)
    if mode == "query":
        ids, d = sharding.search_batch_query_sharded(lambda qs, k: o.search_batch(view, qs, k)[:2], q, 10)
    else:
        def scan(qs, k, kp, b, e):
            ii = np.full((len(qs), k), -1, np.int64); dd = np.full((len(qs), k), np.finfo(np.float32).max, np.float32)
            for i, qq in enumerate(qs):
                a, b_, _, _ = o.exhaustive(view, fab, qq, k, kp, b, e)
                ii[i, :len(a)] = a; dd[i, :len(a)] = b_
            return ii, dd
        ids, d = sharding.exhaustive_search_db_sharded(scan, fab.n, q, 5, 5000, None)
    if rank == 0:
        np.savez(out, ids=ids, d=d)
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["query", "db"])
def test_two_ranks_match_one(tmp_path, mode):
    out = str(tmp_path / "r.npz")
    port = 29500 + (os.getpid() % 500) + (0 if mode == "query" else 500)
    mp.spawn(_worker, args=(2, port, mode, out), nprocs=2, join=True)
    got = np.load(out)
    o = co.Oracle()
    fab = common.fabricate(600, 32, 1, seed=3, layers=1)
    view = o.index_view(fab)
    q = np.random.default_rng(2).standard_normal((21, 32)).astype(np.float32)
    if mode == "query":
        ids, d, _ = o.search_batch(view, q, 10)
        assert np.array_equal(got["ids"], ids) and np.array_equal(got["d"], d)
    else:
        for i in range(len(q)):
            a, b, _, _ = o.exhaustive(view, fab, q[i], 5, 5000)
            assert np.array_equal(got["ids"][i, :len(a)], a.astype(np.int64))
            assert np.array_equal(got["d"][i, :len(a)], b)


def test_device_merge_equals_the_numpy_merge():
    """merge_topk_device (torch ops; the NCCL path's merge) against merge_topk, ties and padding included."""
    import torch

    sys.path.insert(0, str(common.ROOT / "rabitq-ann-search_b200"))
    from cphnsw_b200 import sharding

    rng = np.random.default_rng(0)
    shards, nq, kk = 4, 33, 7
    d = np.sort(rng.integers(0, 50, (shards, nq, kk)).astype(np.float32) * 0.25, axis=2)   # many exact ties
    i = rng.permutation(shards * nq * kk).reshape(shards, nq, kk).astype(np.int64)
    i[1, 3, 4:] = -1; d[1, 3, 4:] = np.finfo(np.float32).max                                  # a short shard list
    i[:, 5, :] = -1; d[:, 5, :] = np.finfo(np.float32).max                                    # a query with no result
    for k in (1, 5, 7, 30):
        wi, wd = sharding.merge_topk(i, d, k)
        gi, gd = sharding.merge_topk_device(torch.from_numpy(i), torch.from_numpy(d), k)
        assert np.array_equal(gi.numpy(), wi) and np.array_equal(gd.numpy().view(np.uint32), wd.view(np.uint32)), k
