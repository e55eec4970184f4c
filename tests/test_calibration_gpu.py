"""GPU: N4, the calibration sample loop (cphnsw_b200_calibration_samples) against the C restatement bit for bit -- and
through it against the composition over the unmodified reference (tests/test_oracle_calibration.py pins the restatement;
tests/golden/calib_golden.npz holds the composition's output on the reference-built index files)."""
import numpy as np
import pytest

import common
from common import co

pytestmark = pytest.mark.gpu
FLOATS = ("nn_dist_sq", "dist_qp_sq", "nop", "ip_corrected", "ip_qo_denom", "true_ip")


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _run(ix, q, start):
    import torch
    from cphnsw_b200 import hooks

    out = hooks.calibration_samples(ix, torch.from_numpy(q), torch.from_numpy(start.view(np.int32)))
    torch.cuda.synchronize()
    r = {k: v.cpu().numpy() for k, v in out.items()}
    r["parent"] = r["parent"].view(np.uint32)
    r["neighbor"] = r["neighbor"].view(np.uint32)
    return r


@pytest.mark.parametrize("dim,bits", [(24, 1), (64, 2), (100, 2), (128, 4), (128, 1), (960, 2), (1500, 4)])
def test_samples_match_the_restatement(oracle, dim, bits):
    fab = common.fabricate(3000, dim, bits, seed=dim + bits, counts=(32, 32, 29, 9, 0))
    ix = common.gpu_index_from(fab)      # (a hole inside a block -- covered on the CPU side -- is refused by the index hand-off as corrupt)
    rng = np.random.default_rng(5)
    ns = 777
    start = rng.integers(0, fab.n, ns).astype(np.uint32)
    q = np.ascontiguousarray(fab.raw[rng.integers(0, fab.n, ns), :dim])
    q[ns // 2:] += (0.2 * rng.standard_normal((ns - ns // 2, dim))).astype(np.float32)
    got = _run(ix, q, start)
    qp = np.zeros((ns, fab.D), np.float32)
    qp[:, :dim] = q
    want = oracle.calibration_samples(oracle.index_view(fab), qp, start)     # coeff_constant as the search path has it: K1's
    for k in ("parent", "neighbor"):
        assert np.array_equal(got[k], want[k]), k
    for k in FLOATS:
        assert np.array_equal(_bits(got[k]), _bits(want[k])), k
    assert (got["parent"] != start).any() and (got["neighbor"] == 0xFFFFFFFF).any()


@pytest.mark.parametrize("bits", [1, 2, 4])
def test_samples_match_the_reference_composition_on_its_own_index_files(bits):
    """Everything bit for bit except ip_corrected, which carries coeff_constant: at D = 32 (these files) GCC fuses that
    coefficient differently inside the composed loop than in the search path K1 reproduces (an ulp of a term ~20 against
    results ~0.1); at D = 128 and 1024 the two contexts agree and the restatement test above is bit-exact on all fields."""
    g = np.load(common.GOLDEN / "calib_golden.npz")
    ix = common.gpu_index_from(common.GOLDEN / f"ref_n300_d24_b{bits}.bin")
    got = _run(ix, g[f"queries_b{bits}"], g[f"start_b{bits}"])
    for k in ("parent", "neighbor"):
        assert np.array_equal(got[k], g[f"{k}_b{bits}"]), k
    for k in ("nn_dist_sq", "dist_qp_sq", "nop", "ip_qo_denom", "true_ip"):
        assert np.array_equal(_bits(got[k]), _bits(g[f"{k}_b{bits}"])), k
    assert np.allclose(got["ip_corrected"], g[f"ip_corrected_b{bits}"], rtol=0, atol=2e-5)


@pytest.mark.skipif(not co.have_ref(), reason="oracle/_ref (compiled reference) not present")
@pytest.mark.parametrize("n,dim,bits", [(20000, 128, 4), (20000, 128, 1)])
def test_samples_match_the_composition_on_a_reference_built_index(oracle, n, dim, bits):
    """A real index (20k x 128, built by the unmodified reference): database vectors and perturbed ones as queries, the start
    vertices a shuffle of the ids as in calibrate_estimator -- all fields bit for bit against the composition over the
    reference's own primitives."""
    path = common.reference_index_file(n, dim, bits)
    sf = co.SaveFile(path)
    ix = common.gpu_index_from(path)
    rng = np.random.default_rng(11)
    ns = 4000
    start = rng.permutation(n)[:ns].astype(np.uint32)
    q = np.ascontiguousarray(sf.raw[rng.permutation(n)[:ns], :dim]).astype(np.float32)
    q[ns // 2:] += (0.5 * rng.standard_normal((ns - ns // 2, dim))).astype(np.float32)
    got = _run(ix, q, start)
    qp = np.zeros((ns, sf.D), np.float32)
    qp[:, :dim] = q
    ref = oracle.ref_calibration_samples(sf, qp, start)
    for k in ("parent", "neighbor"):
        assert np.array_equal(got[k], ref[k]), k
    for k in FLOATS:
        assert np.array_equal(_bits(got[k]), _bits(ref[k])), k
