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
    fab = common.fabricate(6000, 32, 1, seed=3, layers=1)
    view = o.index_view(fab)
    q = np.random.default_rng(2).standard_normal((21, 32)).astype(np.float32)
    if mode == "query":
        ids, d = sharding.search_batch_query_sharded(lambda qs, k: o.search_batch(view, qs, k)[:2], q, 10)
    else:
        import torch

        # CPU stand-ins for the two C-ABI calls of the sharded scan (hooks.exhaustive_candidates / merge_candidates), built
        # on the oracle's estimates and exact distances of this rank's shard; the flow, the piece plan and the collectives
        # are the product's
        B, E = sharding.db_shard(fab.n, rank, world)
        FMAX = np.finfo(np.float32).max
        est, dist_of = [], []
        for qq in q:
            ia, da, _, e_ = o.exhaustive(view, fab, qq, fab.n, fab.n)
            est.append(e_.view(np.uint32).astype(np.uint64))
            dd = np.empty(fab.n, np.float32); dd[ia] = da
            dist_of.append(dd)

        def cand(queries, kp, b, e, off, prior, tau, want):
            keys = np.full((len(q), kp), np.uint64(0xFFFFFFFFFFFFFFFF), np.uint64)
            dists = np.full((len(q), kp), FMAX, np.float32)
            tout = np.full(len(q), FMAX, np.float32)
            for i in range(len(q)):
                loc = np.arange(b, e, dtype=np.uint64)
                kk = (est[i][B + b:B + e] << np.uint64(32)) | loc
                if tau is not None:
                    kk = kk[est[i][B + b:B + e].astype(np.uint32).view(np.float32) <= tau[i].item()]
                if prior is not None:
                    pk = prior[i].numpy().view(np.uint64)
                    kk = np.concatenate([kk, pk[pk != np.uint64(0xFFFFFFFFFFFFFFFF)]])
                kk = np.sort(kk)[:kp]
                if len(kk) == kp:
                    tout[i] = np.uint32(kk[-1] >> np.uint64(32)).view(np.float32)
                if want:
                    dists[i, :len(kk)] = dist_of[i][B + (kk & np.uint64(0xFFFFFFFF)).astype(np.int64)]
                    kk = kk + np.uint64(off)
                keys[i, :len(kk)] = kk
            return torch.from_numpy(keys.view(np.int64)), torch.from_numpy(dists) if want else None, torch.from_numpy(tout)

        def merge(keys, dists, k):
            kk = keys.numpy().view(np.uint64)
            lists, nq, kp = kk.shape
            ids = np.full((nq, max(k, 1)), -1, np.int64); dd = np.full((nq, max(k, 1)), FMAX, np.float32); tout = np.full(nq, FMAX, np.float32)
            for i in range(nq):
                flat = kk[:, i, :].reshape(-1)
                order = np.argsort(flat, kind="stable")[:kp]
                order = order[flat[order] != np.uint64(0xFFFFFFFFFFFFFFFF)]
                if len(order) == kp:
                    tout[i] = np.uint32(flat[order[-1]] >> np.uint64(32)).view(np.float32)
                if dists is not None and k:
                    fd = dists.numpy()[:, i, :].reshape(-1)[order]
                    fi = (flat[order] & np.uint64(0xFFFFFFFF)).astype(np.int64)
                    o2 = np.lexsort((fi, fd))[:k]
                    ids[i, :len(o2)] = fi[o2]; dd[i, :len(o2)] = fd[o2]
            return torch.from_numpy(ids), torch.from_numpy(dd), torch.from_numpy(tout)

        ids, d = sharding.exhaustive_search_db_sharded(cand, merge, E - B, fab.n, B, torch.from_numpy(q), 5, 40, None, prefix=128, growth=2)
        ids, d = ids.numpy(), d.numpy()
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
    fab = common.fabricate(6000, 32, 1, seed=3, layers=1)
    view = o.index_view(fab)
    q = np.random.default_rng(2).standard_normal((21, 32)).astype(np.float32)
    if mode == "query":
        ids, d, _ = o.search_batch(view, q, 10)
        assert np.array_equal(got["ids"], ids) and np.array_equal(got["d"], d)
    else:
        for i in range(len(q)):      # what ONE scan of the whole database returns with the same k' (not with a larger one)
            a, b, _, _ = o.exhaustive(view, fab, q[i], 5, 40)
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
