"""GPU parity: every kernel of the query path against the oracle on identical inputs.

Integer sums, LUT bytes, ids: bit-exact.  Float estimates / bounds / distances: compared bit for bit
as well (the kernels mirror the reference's operation sequences), which is stricter than the 1e-5
relative tolerance BASELINE.json's north_star allows; REL_TOL documents that allowance.
"""
import ctypes as C

import os
import sys

import numpy as np
import pytest

import common
from common import co

pytestmark = pytest.mark.gpu
REL_TOL = 1e-5  # north_star's tolerance for floats; the assertions below demand equality


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _torch():
    import torch

    return torch


# ---------------------------------------------------------------------------------------------------
# K1
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim", [16, 20, 96, 128, 300, 960, 1536])
@pytest.mark.parametrize("center", [False, True])
def test_query_prep_matches_oracle(oracle, dim, center):
    from cphnsw_b200 import hooks

    torch = _torch()
    fab = common.fabricate(40, dim, 1, seed=dim)
    ix = common.gpu_index_from(fab)
    rng = np.random.default_rng(dim)
    q = rng.standard_normal((9, dim)).astype(np.float32)
    q[1] = 0.0                      # all-zero query: delta floor 1e-20
    q[2] = 3.25                     # constant query
    q[3] *= 1e4                     # large dynamic range
    q[4, 1:] = 0.0                  # one-hot
    out = hooks.prepare_queries(ix, torch.from_numpy(q).cuda(), center=center)
    qq = q - fab.centroid[None, :] if center else q
    lut, co_, rot = oracle.encode_queries(qq.astype(np.float32), D=fab.D, want_rotated=True)
    assert np.array_equal(_bits(out["rotated"].cpu().numpy()), _bits(rot))
    assert np.array_equal(out["lut"].cpu().numpy(), lut)
    assert np.array_equal(_bits(out["coeffs"].cpu().numpy()), _bits(co_))
    # bit-planes are the same 4-bit values as the LUT's single-bit entries: u[4j+b] = lut[j][1<<b]
    u = np.stack([lut[:, :, 1 << b] for b in range(4)], axis=2).reshape(q.shape[0], fab.D)
    up = out["uplanes"].cpu().numpy().view(np.uint32)
    W = max(fab.D, 128) // 32
    for t in range(4):
        bits_t = ((u >> t) & 1).astype(np.uint8)
        padded = np.zeros((q.shape[0], W * 32), np.uint8)
        padded[:, :fab.D] = bits_t
        words = np.packbits(padded.reshape(q.shape[0], W, 32), axis=2, bitorder="little").view(np.uint32)[:, :, 0]
        assert np.array_equal(up[:, t, :], words)


# ---------------------------------------------------------------------------------------------------
# K2
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim,bits", [(128, 1), (128, 2), (128, 4), (96, 4), (16, 1), (24, 2), (960, 2), (1024, 4), (2048, 1)])
def test_fastscan_blocks_match_oracle(oracle, dim, bits):
    from cphnsw_b200 import hooks

    torch = _torch()
    n = 96
    fab = common.fabricate(n, dim, bits, seed=7 * dim + bits, counts=(0, 1, 7, 8, 9, 15, 16, 24, 31, 32), degenerate=True,
                           a=1.03, b=-0.01)
    ix = common.gpu_index_from(fab)
    rng = np.random.default_rng(1)
    nq = 5
    q = rng.standard_normal((nq, dim)).astype(np.float32)
    prep = hooks.prepare_queries(ix, torch.from_numpy(q).cuda())
    lut, coeffs = oracle.encode_queries(q, D=fab.D)
    nblocks = 3 * n
    vids = rng.integers(0, n, nblocks).astype(np.uint32)
    qob = rng.integers(0, nq, nblocks).astype(np.uint32)
    dqp = rng.uniform(0.5, 300.0, nblocks).astype(np.float32)
    dqp[::17] = 0.0
    dqp[5::23] = 5e-13
    lvl = rng.integers(0, 6, nblocks).astype(np.int32)
    out = hooks.fastscan_blocks(ix, prep["uplanes"], prep["coeffs"], torch.from_numpy(dqp), vertex_ids=torch.from_numpy(vids.view(np.int32)),
                                query_of_block=torch.from_numpy(qob.view(np.int32)), slack_level=torch.from_numpy(lvl))
    out = {k: v.cpu().numpy() for k, v in out.items()}
    lay, nb = fab.lay, fab.nb_off
    for i in range(nblocks):
        rec = fab.search_data[vids[i]]
        planes = rec[nb:nb + 4 * fab.D * bits]
        f = lambda name, t, cnt: rec[nb + lay[name]:nb + lay[name] + cnt].view(t)  # noqa: E731
        count = int(f("count", np.uint32, 4)[0])
        nbit, msb, msb2 = oracle.fastscan(fab.D, bits, lut[qob[i]], planes)
        assert np.array_equal(out["nbit"][i].view(np.uint32), nbit)
        assert np.array_equal(out["msb"][i].view(np.uint32), msb)
        assert np.array_equal(out["msb2"][i].view(np.uint32), msb2)
        li = min(int(lvl[i]), fab.num_slack_levels - 1)
        params = np.array([*coeffs[qob[i]], fab.affine_a, fab.affine_b, fab.ip_qo_floor, fab.slack_levels[li]], np.float32)
        wpop = f("wpop", np.uint16, 64) if bits > 1 else None
        est, lower, msb_lower = oracle.convert(fab.D, bits, params, nbit, msb, msb2, f("nop", np.float32, 128),
                                               f("ip_qo", np.float32, 128), f("ip_cp", np.float32, 128),
                                               f("pop", np.uint16, 64), wpop, count, float(dqp[i]))
        for name, want in (("est", est), ("lower", lower), ("msb_lower", msb_lower)):
            assert np.array_equal(_bits(out[name][i][:count]), _bits(want[:count])), (name, i, count)


# ---------------------------------------------------------------------------------------------------
# K4 primitive and K3 prologue
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim", [16, 96, 128, 960])
def test_exact_l2_matches_oracle(oracle, dim):
    from cphnsw_b200 import hooks

    torch = _torch()
    fab = common.fabricate(64, dim, 1, seed=dim)
    ix = common.gpu_index_from(fab)
    rng = np.random.default_rng(3)
    q = rng.standard_normal((6, dim)).astype(np.float32)
    ids = rng.integers(0, fab.n, (6, 11)).astype(np.int32)
    got = hooks.exact_l2(ix, torch.from_numpy(q), torch.from_numpy(ids)).cpu().numpy()
    for i in range(6):
        qp = np.zeros(fab.D, np.float32); qp[:dim] = q[i]
        qn = oracle.dot(qp, qp)
        for j in range(11):
            d = oracle.dot(qp, fab.raw[ids[i, j]])
            want = np.float32(np.float32(qn + fab.norm_sq[ids[i, j]]) - np.float32(2.0) * d)
            want = np.float32(0.0) if want < 0 else want
            assert _bits(got[i, j]) == _bits(want)


@pytest.mark.parametrize("dim,layers", [(128, 3), (48, 2), (960, 1)])
def test_greedy_descent_matches_oracle(oracle, dim, layers):
    from cphnsw_b200 import hooks

    torch = _torch()
    fab = common.fabricate(900, dim, 1, seed=layers, layers=layers)
    ix = common.gpu_index_from(fab)
    view = oracle.index_view(fab)
    q = np.random.default_rng(4).standard_normal((40, dim)).astype(np.float32)
    got = hooks.greedy_descent(ix, torch.from_numpy(q)).cpu().numpy().view(np.uint32)
    want = np.array([oracle.greedy_descent(view, q[i]) for i in range(40)], np.uint32)
    assert np.array_equal(got, want)


# ---------------------------------------------------------------------------------------------------
# K3 + K4: the whole query path
# ---------------------------------------------------------------------------------------------------
STAT_KEYS = ("pops", "expansions", "beam_pushes", "nn_pushes", "lb_skips", "gamma_terms", "msb_skipped", "estimated",
             "descent_dists")


def _check_search(oracle, ix, view, q, k, check_stats=True):
    ids, dists = ix.search_batch(q, k)
    st = ix.last_stats()
    oid, od, ost = oracle.search_batch(view, q, k)
    assert ids.shape == (q.shape[0], k) and dists.shape == (q.shape[0], k)
    if k:
        valid = ids >= 0
        d = np.where(valid, dists, np.float32(0))
        assert np.all(np.diff(np.where(valid, dists, np.finfo(np.float32).max), axis=1) >= 0), "rows must ascend"
        gi, gd = common.sorted_rows(ids, dists)
        wi, wd = common.sorted_rows(oid, od)
        assert np.array_equal(gi, wi), f"ids differ in {int((gi != wi).any(axis=1).sum())} of {q.shape[0]} rows"
        assert np.array_equal(_bits(gd), _bits(wd))
        del d
    if check_stats:
        for key in STAT_KEYS:
            assert st[key] == ost[key], (key, st[key], ost[key])
        assert st["max_beam"] == ost["max_beam"]
        assert st["exact_calls"] >= ost["exact_calls"]   # speculative distances are a superset
    return st


@pytest.mark.parametrize("dim,bits", [(128, 1), (128, 2), (128, 4), (96, 4), (20, 1), (960, 2)])
@pytest.mark.parametrize("k", [10, 1, 100])
def test_search_fabricated_graph(oracle, dim, bits, k):
    # small gamma and wide slack so that gamma termination, lower-bound skips and MSB-only skips all fire
    fab = common.fabricate(1500, dim, bits, seed=dim + bits, counts=(32, 32, 32, 31, 24, 9, 0), degenerate=True, layers=2,
                           gamma=1.02, gamma_max=1.6, gamma_beta=0.8, gamma_warmup=4, floor=0.35,
                           slacks=(0.05, 0.08, 0.1, 0.12))
    ix = common.gpu_index_from(fab)
    view = oracle.index_view(fab)
    q = np.random.default_rng(11).standard_normal((64, dim)).astype(np.float32)
    q[0] = fab.raw[3, :dim]   # lands on the duplicated vectors: dist_qp_sq == 0 path
    st = _check_search(oracle, ix, view, q, k)
    assert st["expansions"] > 0


def test_search_exercises_every_branch(oracle):
    tot = dict.fromkeys(("gamma_terms", "lb_skips", "msb_skipped"), 0)
    for seed, gamma in ((1, 1.01), (2, 1.3), (3, 50.0)):
        fab = common.fabricate(2000, 64, 4, seed=seed, counts=(32, 17), gamma=gamma, gamma_max=gamma * 2, gamma_warmup=2,
                               slacks=(0.01, 0.02), floor=0.2)
        ix = common.gpu_index_from(fab)
        q = np.random.default_rng(seed).standard_normal((48, 64)).astype(np.float32)
        st = _check_search(oracle, ix, oracle.index_view(fab), q, 5)
        for key in tot:
            tot[key] += st[key]
    assert tot["gamma_terms"] > 0 and tot["lb_skips"] > 0 and tot["msb_skipped"] > 0, tot


@pytest.mark.parametrize("k", [0, 1, 33, 128, 129, 300, 5000])
def test_search_k_edge_cases(oracle, k):
    fab = common.fabricate(700, 32, 2, seed=k, layers=1, gamma=1e6, gamma_max=1e7)
    ix = common.gpu_index_from(fab)
    q = np.random.default_rng(5).standard_normal((17, 32)).astype(np.float32)
    _check_search(oracle, ix, oracle.index_view(fab), q, k, check_stats=k > 0)
    ids, dists = ix.search_batch(q, k)
    if k > 700:
        assert (ids[:, -1] == -1).any() or (ids >= 0).all()
        pad = ids < 0
        assert np.all(dists[pad] == np.finfo(np.float32).max)


def test_frontier_overflow_is_rerun(oracle):
    fab = common.fabricate(4000, 64, 1, seed=9, gamma=1e6, gamma_max=1e7, slacks=(3.0,), floor=0.2)
    ix = common.gpu_index_from(fab)
    q = np.random.default_rng(6).standard_normal((40, 64)).astype(np.float32)
    view = oracle.index_view(fab)
    st_big = _check_search(oracle, ix, view, q, 10)
    ix.set_option("beam_capacity", 64)
    ids, dists = ix.search_batch(q, 10)
    st = ix.last_stats()
    oid, od, ost = oracle.search_batch(view, q, 10)
    if ost["max_beam"] > 64:
        assert st["overflow_retries"] > 0
    assert st_big["overflow_retries"] == 0
    gi, gd = common.sorted_rows(ids, dists); wi, wd = common.sorted_rows(oid, od)
    assert np.array_equal(gi, wi) and np.array_equal(_bits(gd), _bits(wd))


def test_save_file_loader_equals_upload(oracle, tmp_path):
    fab = common.fabricate(800, 96, 4, seed=21, layers=2, counts=(32, 30))
    path = common.write_save_file(fab, tmp_path / "fab.bin")
    a = common.gpu_index_from(fab)
    b = common.gpu_index_from(path)
    q = np.random.default_rng(8).standard_normal((32, 96)).astype(np.float32)
    ia, da = a.search_batch(q, 10)
    ib, db = b.search_batch(q, 10)
    assert np.array_equal(ia, ib) and np.array_equal(_bits(da), _bits(db))
    _check_search(oracle, b, oracle.index_view(co.SaveFile(path)), q, 10)


def test_api_surface(oracle):
    torch = _torch()
    import cphnsw_b200

    fab = common.fabricate(600, 40, 1, seed=2, layers=1)
    ix = common.gpu_index_from(fab)
    assert ix.size == 600 and ix.dim == 40 and ix.is_finalized
    q = np.random.default_rng(1).standard_normal((8, 40)).astype(np.float32)
    ids, dists = ix.search_batch(q, 7)
    assert ids.dtype == np.int64 and dists.dtype == np.float32
    i0, d0 = ix.search(q[0], 7)
    m = int((ids[0] >= 0).sum())
    assert np.array_equal(i0, ids[0, :m]) and np.array_equal(d0, dists[0, :m])
    i1, _ = ix.search(q[0], 0)          # k=0 searches with k=1 (api/hnsw_index.hpp:187)
    assert i1.shape == (1,)
    tid, td = ix.search_batch(torch.from_numpy(q).cuda(), 7)
    assert tid.is_cuda and np.array_equal(tid.cpu().numpy(), ids) and np.array_equal(td.cpu().numpy(), dists)
    ids64, _ = ix.search_batch(q.astype(np.float64), 7)    # forcecast like pybind's array_t
    assert np.array_equal(ids64, ids)
    with pytest.raises(ValueError, match="queries must be a"):
        ix.search_batch(q[:, :5], 3)
    with pytest.raises(ValueError, match="query must be 1D"):
        ix.search(q, 3)
    with pytest.raises(ValueError, match="Unsupported bits=3"):
        cphnsw_b200.CPIndex(40, 3)
    with pytest.raises(ValueError, match="Unsupported dimension"):
        cphnsw_b200.CPIndex(4096, 1)
    fresh = cphnsw_b200.CPIndex(40, 1)
    with pytest.raises(RuntimeError):
        fresh.search_batch(q, 3)
    with pytest.raises(RuntimeError, match="Finalize called without a pending build"):
        fresh.finalize()
    with pytest.raises(RuntimeError, match="finalized before saving"):
        fresh.save("/tmp/never.bin")
    with pytest.raises(RuntimeError, match="Invalid magic"):
        bad = "/tmp/cphnsw_b200_bad.bin"
        open(bad, "wb").write(b"\0" * 4096)
        fresh.load(bad)


# ---------------------------------------------------------------------------------------------------
# real indexes built by the unmodified reference (oracle/_ref, prebuilt; travels to the GPU box)
# ---------------------------------------------------------------------------------------------------
needs_ref = pytest.mark.skipif(not co.have_ref(), reason="oracle/_ref (compiled reference) not present")


@needs_ref
@pytest.mark.parametrize("n,dim,bits,clusters,k", [(20000, 128, 4, 0, 10), (20000, 128, 1, 0, 10), (20000, 128, 2, 64, 10),
                                                   (6000, 960, 2, 0, 100), (8000, 96, 1, 0, 10)])
def test_search_reference_built_index(oracle, n, dim, bits, clusters, k):
    path = common.reference_index_file(n, dim, bits, clusters)
    sf = co.SaveFile(path)
    ix = common.gpu_index_from(path)
    q = common.queries_for(dim, 200, clusters=clusters)
    _check_search(oracle, ix, oracle.index_view(sf), q, k)
    # and against the reference module itself
    ref = co.ref_module().CPIndex(dim=dim, bits=bits)
    ref.load(str(path))
    rid, rd = ref.search_batch(q, k)
    ids, dists = ix.search_batch(q, k)
    gi, gd = common.sorted_rows(ids, dists); wi, wd = common.sorted_rows(rid, rd)
    assert np.array_equal(gi, wi) and np.array_equal(_bits(gd), _bits(wd))
    # relative tolerance statement of the north star (implied by equality)
    assert np.all(np.abs(gd - wd) <= REL_TOL * np.maximum(np.abs(wd), 1e-30))


@needs_ref
@pytest.mark.parametrize("n,dim,bits,k,nq", [(250_000, 128, 4, 10, 400), (100_000, 960, 2, 100, 200)])
def test_search_reference_built_index_at_scale(n, dim, bits, k, nq):
    """BASELINE configs 2 and 3 at the largest sizes a test run can build (the index comes from the reference's own
    build code through oracle/_ref/refbuild, exactly as bench.py obtains its index -- 250k is past the point where the
    stock calibration stops converging, so this also covers the relaxed-calibration files the bench searches): ids as
    multisets and distance bits against the reference module, for the counting kernels and for the fast ones."""
    import argparse

    sys.path.insert(0, str(common.ROOT))
    import bench

    if not (common.ROOT / "oracle" / "_ref" / "refbuild").exists():
        pytest.skip("oracle/_ref/refbuild not present")
    args = argparse.Namespace(n=n, dim=dim, bits=bits, clusters=0, seed=1234)
    path, _ = bench.obtain_index(args, 0, 1)
    ix = common.gpu_index_from(path)
    q = common.queries_for(dim, nq)
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count())
    ref = co.ref_module().CPIndex(dim=dim, bits=bits)
    ref.load(str(path))
    rid, rd = ref.search_batch(q, k)
    wi, wd = common.sorted_rows(rid, rd)
    for stats in (1, 0):
        ix.set_option("collect_stats", stats)
        ids, dists = ix.search_batch(q, k)
        gi, gd = common.sorted_rows(ids, dists)
        assert np.array_equal(gi, wi), f"ids differ from the reference (collect_stats={stats})"
        assert np.array_equal(_bits(gd), _bits(wd)), f"distance bits differ from the reference (collect_stats={stats})"
    st = ix.last_stats()
    assert st["overflow_retries"] == 0


@needs_ref
def test_drop_in_build_finalize_search(oracle, tmp_path):
    """cphnsw_b200.CPIndex used exactly like cphnsw.CPIndex (build/finalize by the reference module)."""
    import cphnsw_b200

    cphnsw_b200.set_host_module(co.ref_module())
    base = co.synthetic(3000, 64, seed=3)
    q = co.synthetic(50, 64, seed=4)
    mine = cphnsw_b200.CPIndex(64, bits=2)
    mine.build(base)
    mine.finalize()
    mine.save(str(tmp_path / "a.bin"))
    ref = co.ref_module().CPIndex(dim=64, bits=2)
    ref.load(str(tmp_path / "a.bin"))
    rid, rd = ref.search_batch(q, 10)
    ids, dists = mine.search_batch(q, 10)
    gi, gd = common.sorted_rows(ids, dists); wi, wd = common.sorted_rows(rid, rd)
    assert np.array_equal(gi, wi) and np.array_equal(_bits(gd), _bits(wd))
    i0, d0 = mine.search(q[0], 10)
    r0, rd0 = ref.search(q[0], 10)
    assert sorted(i0.tolist()) == sorted(r0.tolist())


# ---------------------------------------------------------------------------------------------------
# committed golden fixtures (written by the unmodified reference, tests/golden/make_golden.py): the CUDA path
# through the C ABI's save-file loader against the reference's own search_batch results -- needs neither the
# oracle nor oracle/_ref at run time
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("bits", [1, 2, 4])
@pytest.mark.parametrize("k", [1, 10, 50])
def test_golden_index_files(bits, k):
    import cphnsw_b200

    g = np.load(common.GOLDEN / "e2e_golden.npz")
    ix = cphnsw_b200.CPIndex(24, bits)
    ix.load(str(common.GOLDEN / f"ref_n300_d24_b{bits}.bin"))
    assert ix.size == 300 and ix.is_finalized
    ids, dists = ix.search_batch(g["queries"], k)
    gi, gd = common.sorted_rows(ids, dists)
    wi, wd = common.sorted_rows(g[f"ids_b{bits}_k{k}"], g[f"dists_b{bits}_k{k}"])
    assert np.array_equal(gi, wi) and np.array_equal(_bits(gd), _bits(wd))


def test_golden_kernel_vectors():
    """K1 / K2 / K4 against the committed reference outputs (k1_golden, k2_golden, l2_golden)."""
    from cphnsw_b200 import hooks

    torch = _torch()
    g1 = np.load(common.GOLDEN / "k1_golden.npz")
    for dim in (16, 20, 96, 128, 960):
        fab = common.fabricate(40, dim, 1, seed=1)
        ix = common.gpu_index_from(fab)
        out = hooks.prepare_queries(ix, torch.from_numpy(g1[f"q_{dim}"]).cuda())
        assert np.array_equal(out["lut"].cpu().numpy(), g1[f"lut_{dim}"])
        assert np.array_equal(_bits(out["coeffs"].cpu().numpy()), _bits(g1[f"coeffs_{dim}"]))
        assert np.array_equal(_bits(out["rotated"].cpu().numpy()), _bits(g1[f"rot_{dim}"]))
    g2 = np.load(common.GOLDEN / "k2_golden.npz")
    for tag in ("128_1", "128_2", "128_4", "960_2", "16_4"):
        dim, bits = map(int, tag.split("_"))
        fab = common.fabricate(10, dim, bits, seed=dim + bits, counts=(32, 31, 24, 9, 8, 1), degenerate=True, a=1.02, b=0.01)
        assert np.array_equal(fab.search_data[:, fab.nb_off:], g2[f"blocks_{tag}"]), "fixture generator drifted"
        ix = common.gpu_index_from(fab)
        # the golden LUTs came from these queries' encodings; recover the bit-planes from the LUT bytes
        lut, coeffs = g2[f"lut_{tag}"], g2[f"coeffs_{tag}"]
        u = np.stack([lut[:, :, 1 << b] for b in range(4)], axis=2).reshape(2, fab.D)
        W = max(fab.D, 128) // 32
        up = np.zeros((2, 4, W), np.uint32)
        for t in range(4):
            pad = np.zeros((2, W * 32), np.uint8); pad[:, :fab.D] = (u >> t) & 1
            up[:, t, :] = np.packbits(pad.reshape(2, W, 32), axis=2, bitorder="little").view(np.uint32)[:, :, 0]
        n = fab.n
        qi = g2[f"qi_{tag}"].astype(np.int32)
        out = hooks.fastscan_blocks(ix, torch.from_numpy(up.view(np.int32)).cuda(), torch.from_numpy(coeffs).cuda(),
                                    torch.from_numpy(g2[f"dqp_{tag}"].astype(np.float32)), vertex_ids=torch.arange(n, dtype=torch.int32),
                                    query_of_block=torch.from_numpy(qi), slack_level=torch.from_numpy((np.arange(n) % 3).astype(np.int32)))
        out = {k_: v.cpu().numpy() for k_, v in out.items()}
        for v in range(n):
            c = int(g2[f"count_{tag}"][v])
            for name in ("nbit", "msb", "msb2"):
                assert np.array_equal(out[name][v].view(np.uint32), g2[f"{name}_{tag}"][v]), (tag, name, v)
            for name in ("est", "lower", "msb_lower"):
                assert np.array_equal(_bits(out[name][v][:c]), _bits(g2[f"{name}_{tag}"][v][:c])), (tag, name, v)


# ---------------------------------------------------------------------------------------------------
# opt-in post-processing: de-duplication and original ids, outside the parity path
# ---------------------------------------------------------------------------------------------------
def _numpy_unique_topk(ids, dists, k, idmap=None):
    out_i = np.full((ids.shape[0], k), -1, np.int64)
    out_d = np.full((ids.shape[0], k), np.finfo(np.float32).max, np.float32)
    for r in range(ids.shape[0]):
        seen, p = set(), 0
        for i, d in zip(ids[r].tolist(), dists[r].tolist()):
            if i < 0 or i in seen or p >= k:
                continue
            seen.add(i)
            out_i[r, p] = i if idmap is None else int(idmap[i])
            out_d[r, p] = d
            p += 1
    return out_i, out_d


@pytest.mark.parametrize("k,ks", [(10, 30), (1, 1), (10, 10), (40, 100), (5, 300)])
def test_unique_topk_is_the_deduplicated_prefix_of_the_parity_result(k, ks):
    fab = common.fabricate(2000, 32, 2, seed=9, layers=1, counts=(32, 31, 7), degenerate=True)
    ix = common.gpu_index_from(fab)
    q = np.random.default_rng(3).standard_normal((37, 32)).astype(np.float32)
    ids, dists = ix.search_batch(q, ks)                      # the reference's convention: duplicates, internal ids
    assert any(len(set(r.tolist())) < len(r) for r in ids) or ks == 1
    ui, ud = ix.search_batch_unique(q, k, k_search=ks)
    wi, wd = _numpy_unique_topk(ids, dists, k)
    assert np.array_equal(ui, wi) and np.array_equal(_bits(ud), _bits(wd))
    for r in ui:
        live = r[r >= 0]
        assert len(set(live.tolist())) == len(live)


@needs_ref
def test_original_ids_and_recall_through_the_drop_in_api():
    """build()/finalize() through the wrapper recover internal -> original ids; with them and de-duplication the
    returned neighbours can be scored against brute force on the caller's own array."""
    import cphnsw_b200

    cphnsw_b200.set_host_module(co.ref_module())
    base = co.synthetic(4000, 48, seed=5, clusters=16)
    q = common.queries_for(48, 64, seed=6, clusters=16, base_seed=5)
    ix = cphnsw_b200.CPIndex(48, bits=4)
    ix.build(base)
    ix.finalize()
    ids, dists = ix.search_batch_unique(q, 10, k_search=40, original_ids=True)
    # distances are those of the returned original rows
    for r in range(len(q)):
        live = ids[r] >= 0
        ex = ((base[ids[r][live]] - q[r]) ** 2).sum(1)
        assert np.allclose(ex, dists[r][live], rtol=1e-4, atol=1e-4)
    gt = np.argsort(((base[None, :, :] - q[:, None, :]) ** 2).sum(2), axis=1)[:, :10]
    recall = np.mean([len(set(a.tolist()) & set(b.tolist())) / 10 for a, b in zip(ids, gt)])
    # (the reference's own search quality on this data; what matters here is that the ids mean something)
    assert recall > 0.7, recall
    # the raw parity result scored the same way is capped by its duplicates
    pi, _ = ix.search_batch(q, 10)
    assert np.mean([len(set(r.tolist())) for r in pi]) < 10
    raw = np.mean([len(set(ix._id_map[a[a >= 0]].tolist()) & set(b.tolist())) / 10 for a, b in zip(pi, gt)])
    assert recall > raw + 0.1, (recall, raw)


@pytest.mark.parametrize("dim,bits", [(128, 4), (128, 1), (64, 2), (960, 2)])
def test_fastscan_streaming_specialisation_equals_the_general_kernel(dim, bits):
    """K2 has a lean instantiation for the streaming case (contiguous blocks, one query, slack level 0, est + lower
    only); it must write exactly what the general kernel writes."""
    from cphnsw_b200 import hooks

    torch = _torch()
    fab = common.fabricate(700, dim, bits, seed=dim + bits, counts=(32, 32, 31, 16, 3, 0), degenerate=True, a=1.02, b=0.01)
    ix = common.gpu_index_from(fab)
    q = torch.from_numpy(np.random.default_rng(4).standard_normal((1, dim)).astype(np.float32)).cuda()
    prep = hooks.prepare_queries(ix, q)
    dqp = torch.from_numpy(np.random.default_rng(5).uniform(0.0, 300.0, fab.n).astype(np.float32))
    lean = hooks.fastscan_blocks(ix, prep["uplanes"], prep["coeffs"], dqp, first_vertex=0, nblocks=fab.n, want=("est", "lower"))
    full = hooks.fastscan_blocks(ix, prep["uplanes"], prep["coeffs"], dqp, first_vertex=0, nblocks=fab.n)
    for name in ("est", "lower"):
        assert np.array_equal(_bits(lean[name].cpu().numpy()), _bits(full[name].cpu().numpy())), name


@pytest.mark.parametrize("dim,bits,k", [(64, 4, 10), (128, 4, 10), (96, 2, 5), (32, 1, 7), (128, 4, 40)])
def test_result_set_follows_the_reference_heap_where_distinct_ids_tie(oracle, dim, bits, k):
    """Forty vectors stored under fifteen ids each: distinct ids of bit-equal distance meet at the result set's eviction boundary
    all the time, and WHICH of them BoundedMaxHeap lets go is a matter of its heap layout (search/rabitq_search.hpp:26-35).  An
    ascending list that evicts its last entry differs from it in a quarter of these rows."""
    fab = common.fabricate(600, dim, bits, seed=7 + dim, counts=(32, 32, 30, 12), layers=1)
    fab.raw[:] = fab.raw[np.arange(600) % 40]
    fab.norm_sq[:] = fab.norm_sq[np.arange(600) % 40]
    ix = common.gpu_index_from(fab)
    view = oracle.index_view(fab)
    q = np.random.default_rng(3).standard_normal((40, dim)).astype(np.float32)
    _check_search(oracle, ix, view, q, k)                      # the counting build
    ix.set_option("collect_stats", 0)                          # the fast one
    ids, dists = ix.search_batch(q, k)
    oid, od, _ = oracle.search_batch(view, q, k)
    assert np.array_equal(ids, oid) and np.array_equal(_bits(dists), _bits(od))   # entry for entry: sort_heap's order of equal distances too


@pytest.mark.skipif(not co.have_ref(), reason="oracle/_ref (compiled reference) not present")
def test_rows_equal_the_reference_modules_where_distinct_ids_tie(tmp_path):
    """The same situation on an index the unmodified reference built and searched itself (sixty vectors stored twenty times
    each): its rows, entry for entry."""
    import cphnsw_b200

    rng = np.random.default_rng(5)
    base = np.tile(rng.standard_normal((60, 32)).astype(np.float32), (20, 1))
    rng.shuffle(base)
    path = tmp_path / "dup.bin"
    ref = co.build_reference_index(base, 4, path, threads=4)
    q = rng.standard_normal((300, 32)).astype(np.float32)
    ix = cphnsw_b200.CPIndex(32, 4)
    ix.load(str(path))
    for k in (10, 3, 25, 40):
        rid, rd = ref.search_batch(q, k)
        ids, dists = ix.search_batch(q, k)
        assert np.array_equal(ids, rid) and np.array_equal(_bits(dists), _bits(rd))
