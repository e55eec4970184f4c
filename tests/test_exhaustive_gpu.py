"""GPU parity of the exhaustive batched scan (K5) against the oracle's composition of reference
primitives (the reference itself has no brute-force mode: SURVEY.md F9 / section 8c)."""
import numpy as np
import pytest

import common
from common import co

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _torch():
    import torch

    return torch


@pytest.mark.parametrize("tc", [2, 1, 0])
@pytest.mark.parametrize("dim,n", [(128, 5000), (96, 3001), (20, 700), (960, 1200), (256, 900)])
def test_estimates_match_oracle(oracle, dim, n, tc):
    from cphnsw_b200 import hooks

    torch = _torch()
    fab = common.fabricate(n, dim, 1, seed=dim, degenerate=True, a=1.01, b=0.003)
    ix = common.gpu_index_from(fab)
    # 2: tcgen05 kind::f16 scan with the screen folded into the contraction (D <= 128), 1: tcgen05 kind::i8 scan
    # (D <= 256), each where it applies; 0: popcount scan
    ix.set_option("exhaustive_tensor_cores", tc)
    view = oracle.index_view(fab)
    q = np.random.default_rng(1).standard_normal((11, dim)).astype(np.float32)
    q[3] = fab.centroid     # |q - c|^2 == 0: the dist_qp_sq < 1e-12 branch
    sums, est = hooks.exhaustive_estimates(ix, torch.from_numpy(q))
    sums, est = sums.cpu().numpy().view(np.uint32), est.cpu().numpy()
    for i in range(q.shape[0]):
        _, _, osums, oest = oracle.exhaustive(view, fab, q[i], 1, 1)
        assert np.array_equal(sums[i], osums)
        assert np.array_equal(_bits(est[i]), _bits(oest))


@pytest.mark.parametrize("tc", [2, 1, 0])
@pytest.mark.parametrize("dim,n,k,kprime", [(128, 6000, 10, 100), (128, 6000, 1, 1), (96, 3001, 10, 1000), (64, 40000, 100, 400),
                                            (960, 1500, 20, 60), (32, 300, 10, 512), (64, 40000, 100, 256), (256, 9000, 10, 30)])
def test_search_matches_oracle(oracle, dim, n, k, kprime, tc):
    from cphnsw_b200 import hooks

    torch = _torch()
    fab = common.fabricate(n, dim, 1, seed=n + k, degenerate=True)
    ix = common.gpu_index_from(fab)
    ix.set_option("exhaustive_tensor_cores", tc)
    view = oracle.index_view(fab)
    q = np.random.default_rng(2).standard_normal((19, dim)).astype(np.float32)
    ids, dists = hooks.exhaustive_search(ix, torch.from_numpy(q), k, kprime)
    ids, dists = ids.cpu().numpy(), dists.cpu().numpy()
    for i in range(q.shape[0]):
        oi, od, _, _ = oracle.exhaustive(view, fab, q[i], k, kprime)
        m = len(oi)
        assert np.array_equal(ids[i, :m], oi.astype(np.int64)), i
        assert np.array_equal(_bits(dists[i, :m]), _bits(od))
        assert np.all(ids[i, m:] == -1) and np.all(dists[i, m:] == np.finfo(np.float32).max)


def test_tensor_core_scan_equals_popcount_scan_on_a_larger_batch(oracle):
    """Several query groups (one of them partial), several vertex slices, thresholds shared between CTAs: the
    tcgen05 scan and the popcount scan must return the same ids and the same distance bits; a sample of the
    queries is checked against the oracle as well."""
    from cphnsw_b200 import hooks

    torch = _torch()
    fab = common.fabricate(150_000, 128, 1, seed=77)
    ix = common.gpu_index_from(fab)
    view = oracle.index_view(fab)
    q = np.random.default_rng(5).standard_normal((700, 128)).astype(np.float32)
    q[5] = fab.centroid
    out = {}
    for tc in (2, 1, 0):
        ix.set_option("exhaustive_tensor_cores", tc)
        ids, dists = hooks.exhaustive_search(ix, torch.from_numpy(q), 10, 100)
        out[tc] = (ids.cpu().numpy(), dists.cpu().numpy())
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(_bits(out[0][1]), _bits(out[1][1]))
    assert np.array_equal(out[0][0], out[2][0]) and np.array_equal(_bits(out[0][1]), _bits(out[2][1]))
    for i in (0, 5, 255, 256, 511, 699):
        oi, od, _, _ = oracle.exhaustive(view, fab, q[i], 10, 100)
        assert np.array_equal(out[1][0][i, :len(oi)], oi.astype(np.int64)), i
        assert np.array_equal(_bits(out[1][1][i, :len(oi)]), _bits(od))


@pytest.mark.parametrize("tc", [2, 1])
@pytest.mark.parametrize("scale", [1e-3, 30.0, 3e3])
def test_screen_survives_data_scale(oracle, tc, scale):
    """The tensor-core forms reject pairs with a screen whose margin is derived from magnitudes (and, in the f16
    form, from power-of-two rescaling of f16 factors): the answer must not depend on the scale of the data.
    Vectors, per-vertex norms and queries scaled together; clustered queries so thresholds get tight."""
    from cphnsw_b200 import hooks

    torch = _torch()
    fab = common.fabricate(30000, 128, 1, seed=11)
    storage = 64
    nop = fab.search_data[:, storage:storage + 4].copy().view(np.float32) * np.float32(scale)
    fab.search_data[:, storage:storage + 4] = nop.view(np.uint8)
    fab.raw *= np.float32(scale)
    fab.norm_sq = (fab.norm_sq * np.float32(scale) * np.float32(scale)).astype(np.float32)
    fab.centroid = (fab.centroid * np.float32(scale)).astype(np.float32)
    ix = common.gpu_index_from(fab)
    ix.set_option("exhaustive_tensor_cores", tc)
    view = oracle.index_view(fab)
    rng = np.random.default_rng(8)
    q = (fab.raw[rng.integers(0, fab.n, 40), :128] + rng.standard_normal((40, 128)).astype(np.float32) * np.float32(0.3 * scale)).astype(np.float32)
    ids, dists = hooks.exhaustive_search(ix, torch.from_numpy(q), 10, 64)
    ids, dists = ids.cpu().numpy(), dists.cpu().numpy()
    for i in range(0, 40, 3):
        oi, od, _, _ = oracle.exhaustive(view, fab, q[i], 10, 64)
        assert np.array_equal(ids[i, :len(oi)], oi.astype(np.int64)), (i, scale)
        assert np.array_equal(_bits(dists[i, :len(oi)]), _bits(od))


@pytest.mark.parametrize("tc", [2, 1, 0])
def test_golden_reference_composition(tc):
    """All three scan forms against tests/golden/exhaustive_golden.npz -- composed from primitives executed by the
    unmodified reference (tests/golden/make_exhaustive_golden.py) on the committed reference-built 1-bit index; needs
    neither the oracle nor oracle/_ref at run time."""
    import cphnsw_b200
    from cphnsw_b200 import hooks

    torch = _torch()
    g = np.load(common.GOLDEN / "exhaustive_golden.npz")
    ix = cphnsw_b200.CPIndex(24, 1)
    ix.load(str(common.GOLDEN / "ref_n300_d24_b1.bin"))
    ix.set_option("exhaustive_tensor_cores", tc)
    q = torch.from_numpy(g["queries"])
    sums, est = hooks.exhaustive_estimates(ix, q)
    sums, est = sums.cpu().numpy().view(np.uint32), est.cpu().numpy()
    for i in range(len(g["queries"])):
        assert np.array_equal(sums[i], g[f"sums_{i}"]), i
        assert np.array_equal(_bits(est[i]), _bits(g[f"est_{i}"])), i
    for k, kp in ((10, 100), (1, 1), (5, 32), (10, 300)):
        ids, dists = hooks.exhaustive_search(ix, q, k, kp)
        ids, dists = ids.cpu().numpy(), dists.cpu().numpy()
        for i in range(len(g["queries"])):
            assert np.array_equal(ids[i], g[f"ids_{i}_k{k}_kp{kp}"]), (i, k, kp)
            assert np.array_equal(_bits(dists[i]), _bits(g[f"dists_{i}_k{k}_kp{kp}"])), (i, k, kp)


def _sharded_on_one_device(ix, q, k, kp, n, world, prefix, growth):
    """exhaustive_search_db_sharded with `world` virtual ranks on one device (the shards are id ranges of one index), the
    collectives done by hand: the same calls, in the same order, as the NCCL path makes."""
    from cphnsw_b200 import hooks, sharding

    torch = _torch()
    shards = [sharding.db_shard(n, r, world) for r in range(world)]
    plans = [sharding.scan_pieces(e - b, world, n, prefix, growth) for b, e in shards]
    assert len({len(p) for p in plans}) == 1, "every rank must make the same number of exchanges"
    keys, dists, tau = [None] * world, [None] * world, None
    for c in range(len(plans[0])):
        last = c == len(plans[0]) - 1
        tl = []
        for r, (B, _) in enumerate(shards):
            pb, pe = plans[r][c]
            keys[r], dists[r], t = hooks.exhaustive_candidates(ix, q, kp, B + pb, B + pe, 0, keys[r], tau, last)
            tl.append(t)
        if last:
            break
        if c == 0:
            tau = torch.stack([sharding.kth_estimate(kk, -(-kp // world)) for kk in keys]).max(0).values      # all-reduce(max)
        else:
            tau = torch.stack(tl).min(0).values                                  # all-reduce(min)
    return hooks.merge_candidates(ix, torch.stack(keys), torch.stack(dists), k)[:2]


@pytest.mark.parametrize("tc", [2, 1, 0])
@pytest.mark.parametrize("dim,n,k,kp,world,prefix,growth", [
    (128, 150_000, 10, 100, 4, 65536, 4),      # the f16 form seeded with the exchanged thresholds, several pieces per shard
    (96, 70_001, 10, 100, 3, 8192, 3),         # ragged shards, many small pieces
    (128, 9_000, 10, 64, 4, 65536, 4),         # shards smaller than the prefix: one piece each
    (64, 30_000, 100, 256, 8, 4096, 2),        # the tensor-core forms' largest k'
    (96, 20_000, 10, 1000, 2, 4096, 4),        # k' beyond them: popcount form
    (960, 3_000, 20, 60, 2, 1024, 2),          # D > 256: popcount form
    (32, 300, 10, 512, 4, 65536, 4),           # fewer vertices than k' in a shard
])
def test_database_shards_merge_to_the_unsharded_answer(dim, n, k, kp, world, prefix, growth, tc):
    """DB-sharded scan = ONE scan of the whole database: the k' best estimates overall, exact re-rank, top-k by (distance,
    id) -- ids and distance bits -- whatever the number of shards, the piece plan and the scan form."""
    from cphnsw_b200 import hooks

    torch = _torch()
    fab = common.fabricate(n, dim, 1, seed=n % 97 + kp, degenerate=True)
    ix = common.gpu_index_from(fab)
    ix.set_option("exhaustive_tensor_cores", tc)
    q = torch.from_numpy(np.random.default_rng(3).standard_normal((37, dim)).astype(np.float32))
    full_i, full_d = hooks.exhaustive_search(ix, q, k, kp, 0, n)
    got_i, got_d = _sharded_on_one_device(ix, q, k, kp, n, world, prefix, growth)
    assert np.array_equal(got_i.cpu().numpy(), full_i.cpu().numpy())
    assert np.array_equal(_bits(got_d.cpu().numpy()), _bits(full_d.cpu().numpy()))


@pytest.mark.parametrize("tc", [2, 1, 0])
@pytest.mark.parametrize("kp", [100, 256])
def test_long_ranges_are_scanned_in_pieces_with_the_single_scan_result(oracle, tc, kp):
    """cphnsw_b200_exhaustive_search on a range past 4 x 64K vertices scans it in pieces (candidates merged and thresholds
    made exact after each): same ids and distance bits as the oracle's single pass; also on a sub-range with an offset."""
    from cphnsw_b200 import hooks

    torch = _torch()
    n = 300_000
    fab = common.fabricate(n, 96, 1, seed=kp, degenerate=True)
    ix = common.gpu_index_from(fab)
    ix.set_option("exhaustive_tensor_cores", tc)
    view = oracle.index_view(fab)
    qn = np.random.default_rng(5).standard_normal((6, 96)).astype(np.float32)
    for b, e in ((0, n), (1234, n - 7)):
        ids, dists = hooks.exhaustive_search(ix, torch.from_numpy(qn), 10, kp, b, e)
        ids, dists = ids.cpu().numpy(), dists.cpu().numpy()
        for i in range(len(qn)):
            oi, od, _, _ = oracle.exhaustive(view, fab, qn[i], 10, kp, b, e)
            assert np.array_equal(ids[i], oi.astype(np.int64))
            assert np.array_equal(_bits(dists[i]), _bits(od))


def test_single_scan_in_pieces_against_the_oracle(oracle):
    """The candidate interface against the oracle directly: keys (estimate bits, ids), thresholds and exact distances."""
    from cphnsw_b200 import hooks

    torch = _torch()
    fab = common.fabricate(5000, 96, 1, seed=8, degenerate=True)
    ix = common.gpu_index_from(fab)
    view = oracle.index_view(fab)
    qn = np.random.default_rng(4).standard_normal((9, 96)).astype(np.float32)
    q = torch.from_numpy(qn)
    kp = 50
    k1, _, t1 = hooks.exhaustive_candidates(ix, q, kp, 0, 2000)
    k2, d2, t2 = hooks.exhaustive_candidates(ix, q, kp, 2000, 5000, 7, k1, t1, True)
    k2, d2, t1, t2 = k2.cpu().numpy(), d2.cpu().numpy(), t1.cpu().numpy(), t2.cpu().numpy()
    for i in range(len(qn)):
        ids_all, dist_all, _, est = oracle.exhaustive(view, fab, qn[i], 5000, 5000)
        dist_of = dict(zip(ids_all.tolist(), dist_all.tolist()))
        order = np.lexsort((np.arange(5000), _bits(est)))[:kp]
        want_keys = (_bits(est)[order].astype(np.uint64) << np.uint64(32)) | (order.astype(np.uint64) + np.uint64(7))
        assert np.array_equal(k2[i].view(np.uint64), want_keys)
        assert np.array_equal(_bits(d2[i]), _bits(np.float32([dist_of[int(v)] for v in order])))
        first = np.sort(_bits(est[:2000]))[kp - 1]
        assert _bits(t1[i:i + 1])[0] == first and _bits(t2[i:i + 1])[0] == _bits(est)[order[-1]]


def test_rejects_what_it_cannot_do(oracle):
    from cphnsw_b200 import hooks

    torch = _torch()
    ix4 = common.gpu_index_from(common.fabricate(400, 64, 4, seed=1))
    with pytest.raises(ValueError, match="bits=1"):
        hooks.exhaustive_search(ix4, torch.zeros(2, 64), 5, 10)
    ix1 = common.gpu_index_from(common.fabricate(400, 64, 1, seed=1))
    with pytest.raises(ValueError, match="kprime"):
        hooks.exhaustive_search(ix1, torch.zeros(2, 64), 5, 5000)
    with pytest.raises(ValueError, match="range"):
        hooks.exhaustive_search(ix1, torch.zeros(2, 64), 5, 10, 0, 401)


needs_ref = pytest.mark.skipif(not co.have_ref(), reason="oracle/_ref (compiled reference) not present")


@needs_ref
def test_on_a_reference_built_index(oracle):
    from cphnsw_b200 import hooks

    torch = _torch()
    path = common.reference_index_file(8000, 96, 1)
    sf = co.SaveFile(path)
    ix = common.gpu_index_from(path)
    view = oracle.index_view(sf)
    q = common.queries_for(96, 12)
    ids, dists = hooks.exhaustive_search(ix, torch.from_numpy(q), 10, 200)
    ids, dists = ids.cpu().numpy(), dists.cpu().numpy()
    base = np.asarray(sf.raw[:, :96])
    hits = 0
    for i in range(len(q)):
        oi, od, _, _ = oracle.exhaustive(view, sf, q[i], 10, 200)
        assert np.array_equal(ids[i], oi.astype(np.int64)) and np.array_equal(_bits(dists[i]), _bits(od))
        gt = np.argsort(((base - q[i]) ** 2).sum(1))[:10]
        hits += len(set(gt.tolist()) & set(ids[i].tolist()))
    assert hits / (10 * len(q)) > 0.9   # 1-bit estimate + rerank depth 200 finds the true neighbours
