"""GPU parity of the build-side neighbour-code encoder (N3) through the C ABI (cphnsw_b200_neighbor_codes)
against the oracle's restatement of compute_neighbor_aux / compute_neighbor_aux_nbit (pinned to the compiled reference in
tests/test_oracle_build_side.py), and against the neighbour blocks of indexes built and saved by the reference itself.
The same kernel source runs on the CPU in tests/test_neighbor_codes_emulated.py."""
import subprocess

import numpy as np
import pytest

import common
from common import co

pytestmark = pytest.mark.gpu


def _torch():
    import torch

    return torch


def _ids(a):
    return _torch().from_numpy(np.ascontiguousarray(a, np.uint32).view(np.int32))


def _bits(t):
    return t.cpu().numpy().view(np.uint32)


@pytest.mark.parametrize("tile", [0, 1])
@pytest.mark.parametrize("bits", [1, 2, 4])
@pytest.mark.parametrize("dim", [10, 20, 64, 96, 128, 300, 960, 1500])
def test_neighbor_codes_match_the_oracle(oracle, dim, bits, tile):
    import cphnsw_b200
    from cphnsw_b200 import hooks

    torch = _torch()
    npar = 12 if dim <= 128 else 3
    vec, pids, nbr = common.neighbor_code_case(dim, npar, 100 * dim + bits)
    want_c, want_a = common.expected_neighbor_codes(oracle, dim, bits, vec, pids, nbr)
    ix = cphnsw_b200.CPIndex(dim, bits)                  # no index data: the build side runs before one exists
    ix.set_option("neighbor_codes_tile", tile)           # 0: per-warp tiles in global memory / L2 (default), 1: in shared memory
    codes, aux, blocks = hooks.neighbor_codes(ix, torch.from_numpy(vec), _ids(nbr), _ids(pids), blocks=True)
    assert np.array_equal(codes.cpu().numpy(), want_c)
    assert np.array_equal(_bits(aux), want_a.view(np.uint32))
    lay = co.nb_layout(max(16, 1 << (dim - 1).bit_length()), bits)
    want_b = common.blocks_from_codes(dim, bits, want_c, want_a, nbr, len(vec))
    assert np.array_equal(blocks.cpu().numpy()[:, :lay["count"] + 4], want_b[:, :lay["count"] + 4])


@pytest.mark.parametrize("bits", [1, 2, 4])
def test_neighbor_blocks_of_a_reference_built_index_are_regenerated(bits):
    """From the raw vectors and neighbour ids of an index built and saved by the unmodified reference: every neighbour
    block, bit for bit up to its count (the slots past it hold leftovers of earlier refinement rounds)."""
    import cphnsw_b200
    from cphnsw_b200 import hooks

    torch = _torch()
    sf = co.SaveFile(common.GOLDEN / f"ref_n300_d24_b{bits}.bin")
    ids, cnt = sf.field("ids").copy(), sf.field("count")
    for p in range(sf.n):
        ids[p, cnt[p]:] = 0xFFFFFFFF
    ix = cphnsw_b200.CPIndex(sf.dim, bits)
    _, _, blocks = hooks.neighbor_codes(ix, torch.from_numpy(np.ascontiguousarray(sf.raw[:, :sf.dim])), _ids(ids),
                                        rotation_seed=sf.rotation_seed, blocks=True)
    blocks = blocks.cpu().numpy()
    lay, ref = sf.lay, sf.search_data[:, sf.nb_off:]
    for p in range(sf.n):
        c = int(cnt[p])
        got_pl = blocks[p, :4 * sf.D * bits].reshape(bits, sf.D // 8, 32)
        ref_pl = ref[p, :4 * sf.D * bits].reshape(bits, sf.D // 8, 32)
        assert np.array_equal(got_pl[:, :, :c], ref_pl[:, :, :c]), p
        for name, width in (("nop", 4), ("ip_qo", 4), ("ip_cp", 4), ("pop", 2), ("wpop", 2), ("ids", 4)):
            if lay[name] is not None:
                assert np.array_equal(blocks[p, lay[name]:lay[name] + width * c], ref[p, lay[name]:lay[name] + width * c]), (p, name)
        assert blocks[p, lay["count"]:lay["count"] + 4].view(np.uint32)[0] == c


def test_neighbor_codes_argument_errors():
    import cphnsw_b200
    from cphnsw_b200 import _capi

    torch = _torch()
    ix = cphnsw_b200.CPIndex(32, 2)
    v = torch.zeros((4, 32), device="cuda")
    nb = torch.zeros((1, 32), dtype=torch.int32, device="cuda")
    out = torch.zeros(4096, dtype=torch.uint8, device="cuda")
    call = ix._lib.cphnsw_b200_neighbor_codes
    assert call(ix.handle, 32, 3, 42, v.data_ptr(), 32, 4, 0, nb.data_ptr(), 1, out.data_ptr(), 0, 0, 0, 0) == _capi.EINVAL   # bits
    assert call(ix.handle, 32, 2, 42, v.data_ptr(), 16, 4, 0, nb.data_ptr(), 1, out.data_ptr(), 0, 0, 0, 0) == _capi.EINVAL   # stride < dim
    assert call(ix.handle, 32, 2, 42, v.data_ptr(), 32, 4, 0, nb.data_ptr(), 1, 0, 0, 0, 0, 0) == _capi.EINVAL               # no output
    assert call(ix.handle, 32, 2, 42, v.data_ptr(), 32, 4, 0, nb.data_ptr(), 1, 0, 0, out.data_ptr(), 64, 0) == _capi.EINVAL  # block stride too small
    assert call(ix.handle, 32, 2, 42, v.data_ptr(), 32, 4, 0, nb.data_ptr(), 0, 0, 0, 0, 0, 0) == 0                          # nothing to do


def test_standalone_checker_passes():
    """tests/native/neighbor_codes_gpu_check.cpp: the same comparison without Python, plus timing samples."""
    exe = common.ROOT / "tests" / "native" / "_build" / "neighbor_codes_gpu_check"
    exe.parent.mkdir(exist_ok=True)
    root = common.ROOT
    cmd = ["g++", "-O1", "-std=c++17", "-I", str(root / "include"), "-I", str(root / "oracle"), "-I", "/usr/local/cuda/include",
           str(root / "tests" / "native" / "neighbor_codes_gpu_check.cpp"),
           str(root / "rabitq-ann-search_b200" / "cphnsw_b200" / "libcphnsw_b200.so"), str(co.build_port()),
           "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + str(root / "rabitq-ann-search_b200" / "cphnsw_b200"),
           "-Wl,-rpath," + str(root / "oracle"), "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-2000:]
