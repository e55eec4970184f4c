"""CPU: the calibration sample loop (N4; Index::calibrate_estimator, api/hnsw_index.hpp:786-866).  The C restatement
(cpo_calibration_sample) against the golden vectors written by the composition over the unmodified reference
(tests/golden/make_calib_golden.py), and -- where oracle/_ref is present -- against that composition live, on random
samples of the reference-built index files and of fabricated indexes with partial blocks."""
import numpy as np
import pytest

import common
from common import co

FLOATS = ("nn_dist_sq", "dist_qp_sq", "nop", "ip_corrected", "ip_qo_denom", "true_ip")
INTS = ("parent", "neighbor")


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _pad(q, D):
    out = np.zeros((len(q), D), np.float32)
    out[:, :q.shape[1]] = q
    return out


@pytest.fixture()
def shim_context(oracle):
    """coeff_constant as GCC contracts it inside the composed loop (cpo_encode_query: variant 1); restored afterwards."""
    oracle.lib.cpo_set_encode_variant(1)
    yield oracle
    oracle.lib.cpo_set_encode_variant(0)


@pytest.mark.parametrize("bits", [1, 2, 4])
def test_restatement_reproduces_the_reference_composition(shim_context, bits):
    o = shim_context
    g = np.load(common.GOLDEN / "calib_golden.npz")
    sf = co.SaveFile(common.GOLDEN / f"ref_n300_d24_b{bits}.bin")
    got = o.calibration_samples(o.index_view(sf), _pad(g[f"queries_b{bits}"], sf.D), g[f"start_b{bits}"])
    for k in INTS:
        assert np.array_equal(got[k], g[f"{k}_b{bits}"]), k
    for k in FLOATS:
        assert np.array_equal(_bits(got[k]), _bits(g[f"{k}_b{bits}"])), k
    # the loop did something: parents moved off the start vertex, blocks are full, residuals are small
    assert (got["parent"] != g[f"start_b{bits}"]).any() and (got["neighbor"] != 0xFFFFFFFF).all()
    est = got["ip_corrected"] / got["ip_qo_denom"]
    assert np.median(np.abs(est - got["true_ip"])) < 0.5


def test_the_contraction_found_is_the_only_one_that_fits(shim_context):
    """flags = 3 (ip_approx = fma(A', fs, Bc' pc) + C; true_ip with a separate multiply and add) is what the compiled
    composition does; every other candidate differs somewhere -- so the restatement's float sequences are pinned, not assumed."""
    o = shim_context
    g = np.load(common.GOLDEN / "calib_golden.npz")
    sf = co.SaveFile(common.GOLDEN / "ref_n300_d24_b4.bin")
    view = o.index_view(sf)
    q, st = _pad(g["queries_b4"], sf.D), g["start_b4"]
    fits = []
    for flags in range(8):
        got = o.calibration_samples(view, q, st, flags)
        fits.append(all(np.array_equal(_bits(got[k]), _bits(g[f"{k}_b4"])) for k in FLOATS))
    assert fits == [f == 3 for f in range(8)]
    assert o.CALIB_FLAGS == 3
    # and the coefficient context matters: with the search path's coeff_constant (variant 0) ip_corrected moves by an ulp of
    # its large terms, nothing else moves
    o.lib.cpo_set_encode_variant(0)
    got = o.calibration_samples(view, q, st)
    assert not np.array_equal(_bits(got["ip_corrected"]), _bits(g["ip_corrected_b4"]))
    assert np.allclose(got["ip_corrected"], g["ip_corrected_b4"], rtol=0, atol=2e-5)
    for k in ("nn_dist_sq", "dist_qp_sq", "nop", "ip_qo_denom", "true_ip"):
        assert np.array_equal(_bits(got[k]), _bits(g[f"{k}_b4"])), k


needs_ref = pytest.mark.skipif(not co.have_ref(), reason="oracle/_ref (compiled reference) not present")


def encode_variant_for(D):
    """Which way GCC fuses coeff_constant inside the composed loop depends on the template instantiation: the D = 32 one
    comes out as variant 1, D = 128 and D = 1024 (the shapes of every BASELINE config) as variant 0 -- the search path's."""
    return 1 if D == 32 else 0


@needs_ref
@pytest.mark.parametrize("dim,bits", [(24, 1), (100, 2), (128, 4), (128, 1), (960, 2), (1000, 4)])
def test_restatement_equals_the_composition_on_fabricated_indexes(oracle, dim, bits):
    """Partial blocks (count 29, 9, 0) and a hole inside a block (the loop stops at the first empty slot); and the float
    sequences are pinned by enumeration: of the 2 x 8 candidate contractions exactly one reproduces the compiled composition."""
    o = oracle
    fab = common.fabricate(400, dim, bits, seed=dim + bits, counts=(32, 32, 29, 9, 0))
    ids_off = fab.nb_off + co.nb_layout(fab.D, bits)["ids"]
    fab.search_data[7, ids_off + 4 * 5:ids_off + 4 * 6] = 0xFF        # a hole at slot 5 of vertex 7
    rng = np.random.default_rng(3)
    ns = 120
    start = rng.integers(0, 400, ns).astype(np.uint32)
    start[:4] = 7
    q = fab.raw[rng.integers(0, 400, ns)].copy()
    q[:, :dim] += (0.2 * rng.standard_normal((ns, dim))).astype(np.float32)
    ref = o.ref_calibration_samples(fab, q, start)
    view = o.index_view(fab)
    fits = []
    try:
        for ev in (0, 1):
            o.lib.cpo_set_encode_variant(ev)
            for flags in range(8):
                got = o.calibration_samples(view, q, start, flags)
                for k in INTS:
                    assert np.array_equal(got[k], ref[k]), k
                if all(np.array_equal(_bits(got[k]), _bits(ref[k])) for k in FLOATS):
                    fits.append((ev, flags))
    finally:
        o.lib.cpo_set_encode_variant(0)
    assert fits == [(encode_variant_for(fab.D), o.CALIB_FLAGS)]
    assert (ref["neighbor"] == 0xFFFFFFFF).any()
