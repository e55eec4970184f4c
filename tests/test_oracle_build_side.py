"""Build-side oracle groundwork: the neighbour-code encoder of prune_and_write
(graph/graph_refinement.hpp:46-67 -> RaBitQEncoder::compute_neighbor_aux, encoder/rabitq_encoder.hpp:138-181).
No CUDA counterpart exists yet; this pins the C restatements of the 1-bit and N-bit (CAQ) encoders against the unmodified reference
(live where oracle/_ref exists, and through a committed fixture written by it) so the kernel can be built against it."""
import numpy as np
import pytest

import common
from common import co

needs_ref = pytest.mark.skipif(not co.have_ref(), reason="oracle/_ref (compiled reference) not present")


def _cases(dim, n, seed):
    rng = np.random.default_rng(seed)
    parent = rng.standard_normal(dim).astype(np.float32)
    nbrs = (parent + rng.standard_normal((n, dim)).astype(np.float32) * rng.uniform(0.05, 3.0, (n, 1)).astype(np.float32)).astype(np.float32)
    nbrs[0] = parent                      # nop == 0: the norm_epsilon branch
    return parent, nbrs


@needs_ref
@pytest.mark.parametrize("dim", [64, 96, 128, 300, 960])
def test_neighbor_aux_1bit_restatement_equals_the_reference(oracle, dim):
    parent, nbrs = _cases(dim, 200, dim)
    rc, ra = oracle.neighbor_aux(dim, 1, parent, nbrs, ref=True)
    oc, oa = oracle.neighbor_aux(dim, 1, parent, nbrs, ref=False, fused=0)
    assert np.array_equal(rc, oc)
    assert np.array_equal(ra.view(np.uint32), oa.view(np.uint32))
    # the other contraction of `nop_sq += d * d` is NOT what the reference's compiler emitted
    _, ua = oracle.neighbor_aux(dim, 1, parent, nbrs, ref=False, fused=1)
    assert not np.array_equal(ra.view(np.uint32), ua.view(np.uint32))


def test_neighbor_aux_1bit_against_the_committed_fixture(oracle):
    g = np.load(common.GOLDEN / "nbaux_golden.npz")
    for dim in (96, 128):
        oc, oa = oracle.neighbor_aux(dim, 1, g[f"parent_{dim}"], g[f"nbrs_{dim}"], ref=False, fused=0)
        assert np.array_equal(oc, g[f"codes_{dim}_b1"])
        assert np.array_equal(oa.view(np.uint32), g[f"aux_{dim}_b1"].view(np.uint32))


@needs_ref
@pytest.mark.parametrize("B", [2, 4])
@pytest.mark.parametrize("dim", [64, 96, 128, 300, 960])
def test_neighbor_aux_nbit_restatement_equals_the_reference(oracle, dim, B):
    parent, nbrs = _cases(dim, 200 if dim < 900 else 60, dim + B)
    rc, ra = oracle.neighbor_aux(dim, B, parent, nbrs, ref=True)
    oc, oa = oracle.neighbor_aux(dim, B, parent, nbrs)
    assert np.array_equal(rc, oc)
    assert np.array_equal(ra.view(np.uint32), oa.view(np.uint32))


@needs_ref
def test_nbit_contraction_bits_the_search_decides(oracle):
    """Flipping any of the decided bits (1, 2, 3, 5, 7, 8, 9; cphnsw_oracle.c) moves some output off the reference's."""
    parent, nbrs = _cases(128, 3000, 77)
    rc, ra = oracle.neighbor_aux(128, 4, parent, nbrs, ref=True)
    for bit in (1, 2, 3, 5, 7, 8, 9):
        uc, ua = oracle.neighbor_aux(128, 4, parent, nbrs, fused=oracle.CAQ_FLAGS ^ (1 << bit))
        assert not (np.array_equal(rc, uc) and np.array_equal(ra.view(np.uint32), ua.view(np.uint32))), bit


@pytest.mark.parametrize("B", [2, 4])
def test_neighbor_aux_nbit_against_the_committed_fixture(oracle, B):
    g = np.load(common.GOLDEN / "nbaux_golden.npz")
    for dim in (96, 128):
        oc, oa = oracle.neighbor_aux(dim, B, g[f"parent_{dim}"], g[f"nbrs_{dim}"])
        assert np.array_equal(oc, g[f"codes_{dim}_b{B}"])
        assert np.array_equal(oa.view(np.uint32), g[f"aux_{dim}_b{B}"].view(np.uint32))


@needs_ref
def test_fixture_is_what_the_reference_writes(oracle):
    g = np.load(common.GOLDEN / "nbaux_golden.npz")
    for dim in (96, 128):
        for B in (1, 2, 4):
            rc, ra = oracle.neighbor_aux(dim, B, g[f"parent_{dim}"], g[f"nbrs_{dim}"], ref=True)
            assert np.array_equal(rc, g[f"codes_{dim}_b{B}"]) and np.array_equal(ra.view(np.uint32), g[f"aux_{dim}_b{B}"].view(np.uint32))


def _check_against_index_file(oracle, path):
    sf = co.SaveFile(path)
    planes, ids, cnt = sf.field("planes"), sf.field("ids"), sf.field("count")
    nop, ip_qo, ip_cp = sf.field("nop"), sf.field("ip_qo"), sf.field("ip_cp")
    pairs = 0
    for p in range(sf.n):
        c = int(cnt[p])
        if c == 0:
            continue
        oc, oa = oracle.neighbor_aux(sf.dim, sf.B, sf.raw[p][:sf.dim], sf.raw[ids[p][:c]][:, :sf.dim])
        stored = np.transpose(planes[p].reshape(sf.B, sf.D // 8, 32), (2, 0, 1))[:c]     # packed[sp][v] = byte sp of slot v
        assert np.array_equal(stored, oc), p
        aux = np.stack([nop[p][:c], ip_qo[p][:c], ip_cp[p][:c]], 1)
        assert np.array_equal(aux.view(np.uint32), oa.view(np.uint32)), p
        pairs += c
    return pairs


@pytest.mark.parametrize("bits", [1, 2, 4])
def test_restatement_reproduces_the_blocks_of_the_golden_index_files(oracle, bits):
    """What prune_and_write stored in indexes built and saved by the unmodified reference (D = 32; committed files)."""
    assert _check_against_index_file(oracle, common.GOLDEN / f"ref_n300_d24_b{bits}.bin") == 9600


@needs_ref
@pytest.mark.parametrize("n,dim,bits", [(2000, 128, 4), (1500, 96, 2), (1000, 64, 1), (600, 960, 2)])
def test_restatement_reproduces_the_blocks_of_indexes_built_by_the_reference_module(oracle, n, dim, bits):
    """The same against the compiled pybind module (oracle/_ref), whose own instantiation and inlining of the encoder --
    not the shim's -- wrote these blocks: the contraction pattern pinned through the shim holds there too."""
    assert _check_against_index_file(oracle, common.reference_index_file(n, dim, bits)) > 10000
