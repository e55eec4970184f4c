"""GPU, world_size 2 over NCCL (skipped on a 1-GPU box): both multi-GPU modes against the single-GPU
answer -- query-sharded graph search (index replicated, no data-path collective) and DB-sharded
exhaustive scan (per-shard candidates, thresholds exchanged by all-reduce, NCCL all-gather, merge) -- which must equal ONE
scan of the whole database."""
import os
import sys

import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu


def _ngpus():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:  # noqa: BLE001
        return 0


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    sys.path.insert(0, str(common.ROOT / "rabitq-ann-search_b200"))
    import cphnsw_b200
    from cphnsw_b200 import hooks, sharding

    fab4 = common.fabricate(3000, 64, 4, seed=5, layers=2)
    fab1 = common.fabricate(5000, 96, 1, seed=6)
    q4 = np.random.default_rng(1).standard_normal((101, 64)).astype(np.float32)
    q1 = np.random.default_rng(2).standard_normal((23, 96)).astype(np.float32)

    def make(fab):
        ix = cphnsw_b200.CPIndex(fab.dim, fab.B, device=rank)
        hooks.upload_arrays(ix, D=fab.D, bits=fab.B, dim=fab.dim, search_data=fab.search_data, raw=fab.raw, norm_sq=fab.norm_sq,
                            calibration=fab.calibration, centroid=fab.centroid, max_level=fab.max_level, entry_point=fab.entry_point,
                            graph_entry_point=fab.entry_point, rotation_seed=fab.rotation_seed, layers=fab.layers)
        return ix

    ix4, ix1 = make(fab4), make(fab1)
    gi, gd = sharding.search_batch_query_sharded(lambda qs, k: ix4.search_batch(qs, k), q4, 10)

    # DB-sharded scan: every rank holds the whole index here and scans only its id range (ids are global already)
    B, E = sharding.db_shard(fab1.n, rank, world)
    q1d = torch.from_numpy(q1).cuda()

    def cand(qs, kp, b, e, off, prior, tau, want):
        return hooks.exhaustive_candidates(ix1, qs, kp, B + b, B + e, 0, prior, tau, want)

    def merge(keys, dists, k):
        return hooks.merge_candidates(ix1, keys, dists, k)

    ei, ed = sharding.exhaustive_search_db_sharded(cand, merge, E - B, fab1.n, 0, q1d, 10, 64, None, prefix=2048, growth=2)
    ei, ed = ei.cpu().numpy(), ed.cpu().numpy()
    if rank == 0:
        si, sd = ix4.search_batch(q4, 10)
        mi, md = hooks.exhaustive_search(ix1, q1d, 10, 64, 0, fab1.n)       # ONE scan of the whole database
        np.savez(out, gi=gi, gd=gd, si=si, sd=sd, ei=ei, ed=ed, mi=mi.cpu().numpy(), md=md.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpus() < 2, reason="needs 2 GPUs")
def test_two_gpus_match_one(tmp_path):
    import torch.multiprocessing as mp

    out = str(tmp_path / "r.npz")
    mp.spawn(_worker, args=(2, 29731 + os.getpid() % 200, out), nprocs=2, join=True)
    r = np.load(out)
    assert np.array_equal(r["gi"], r["si"]) and np.array_equal(r["gd"].view(np.uint32), r["sd"].view(np.uint32))
    assert np.array_equal(r["ei"], r["mi"]) and np.array_equal(r["ed"].view(np.uint32), r["md"].view(np.uint32))
