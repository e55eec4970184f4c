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
