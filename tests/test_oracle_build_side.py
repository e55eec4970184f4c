"""Build-side oracle groundwork: the neighbour-code encoder of prune_and_write
(graph/graph_refinement.hpp:46-67 -> RaBitQEncoder::compute_neighbor_aux, encoder/rabitq_encoder.hpp:138-181).
No CUDA counterpart exists yet; this pins the C restatement of the 1-bit encoder against the unmodified reference
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
def test_fixture_is_what_the_reference_writes(oracle):
    """The committed fixture also carries the reference's 2- and 4-bit codes (CAQ quantiser) for the next round."""
    g = np.load(common.GOLDEN / "nbaux_golden.npz")
    for dim in (96, 128):
        for B in (1, 2, 4):
            rc, ra = oracle.neighbor_aux(dim, B, g[f"parent_{dim}"], g[f"nbrs_{dim}"], ref=True)
            assert np.array_equal(rc, g[f"codes_{dim}_b{B}"]) and np.array_equal(ra.view(np.uint32), g[f"aux_{dim}_b{B}"].view(np.uint32))
