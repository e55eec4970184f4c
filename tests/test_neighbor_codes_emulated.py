"""N3 kernel logic on the CPU: rabitq-ann-search_b200/csrc/neighbor_codes.cu compiled for the host over
tests/native/cuda_emul.h (one thread per lane, barriers for the warp primitives, IEEE operations for the _rn
intrinsics) and compared bit for bit with the oracle -- the same source the GPU runs, minus the launch.  The GPU
run of the same cases is tests/test_neighbor_codes_gpu.py."""
import ctypes as C
import subprocess

import numpy as np
import pytest

import common
from common import co

NATIVE = common.ROOT / "tests" / "native"


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    out = tmp_path_factory.mktemp("emul") / "libneighbor_codes_emul.so"
    cmd = ["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-I/usr/local/cuda/include",
           str(NATIVE / "neighbor_codes_emul.cpp"), "-o", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return C.CDLL(str(out))


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _run(emul, oracle, dim, bits, vec, pids, nbr, max_warps=0, stride=None, want_blocks=False, global_tile=False):
    D = max(16, 1 << (dim - 1).bit_length())
    stride = stride or dim
    rows = np.zeros((len(vec), stride), np.float32)
    rows[:, :dim] = vec
    rows[:, dim:] = 7.0                                  # what lies beyond dim in a row is not the vector
    signs = oracle.rotation_signs(D)
    codes = np.full((len(nbr), 32, bits, D // 8), 0xAA, np.uint8)
    aux = np.full((len(nbr), 32, 3), np.nan, np.float32)
    bsize = co.nb_layout(D, bits)["size"]
    blocks = np.full((len(nbr), bsize), 0xCC, np.uint8) if want_blocks else None
    used = C.c_uint32(0)
    rc = emul.emul_neighbor_codes(C.c_uint32(dim), C.c_uint32(bits), _p(signs, C.c_float), _p(rows, C.c_float), C.c_uint64(stride),
                                  C.c_uint64(len(vec)), None if pids is None else _p(pids, C.c_uint32), _p(nbr, C.c_uint32),
                                  C.c_uint64(len(nbr)), _p(codes, C.c_uint8), _p(aux, C.c_float),
                                  None if blocks is None else _p(blocks, C.c_uint8), C.c_uint64(bsize), C.c_uint32(max_warps | (0x10000 if global_tile else 0)),
                                  C.byref(used))
    assert rc == 0
    return (codes, aux, used.value, blocks) if want_blocks else (codes, aux, used.value)


@pytest.mark.parametrize("bits", [1, 2, 4])
@pytest.mark.parametrize("dim", [10, 20, 64, 96, 128, 300])
def test_emulated_kernel_equals_the_oracle(emul, oracle, dim, bits):
    vec, pids, nbr = common.neighbor_code_case(dim, 12, 100 * dim + bits)
    want_c, want_a = common.expected_neighbor_codes(oracle, dim, bits, vec, pids, nbr)
    codes, aux, rows = _run(emul, oracle, dim, bits, vec, pids, nbr)
    assert rows == 32
    assert np.array_equal(codes, want_c)
    assert np.array_equal(aux.view(np.uint32), want_a.view(np.uint32))


@pytest.mark.parametrize("bits", [1, 2, 4])
@pytest.mark.parametrize("dim,npar", [(20, 40), (128, 37), (300, 18), (960, 3)])
def test_emulated_kernel_with_tiles_in_global_memory(emul, oracle, dim, bits, npar):
    """The large-D mode: [coordinate][lane] tiles in a scratch buffer, a bounded grid of warps looping over the parents
    (the emulated device has 2 SMs, so every warp takes several parents here)."""
    vec, pids, nbr = common.neighbor_code_case(dim, npar, 7 * dim + bits)
    want_c, want_a = common.expected_neighbor_codes(oracle, dim, bits, vec, pids, nbr)
    codes, aux, _, blocks = _run(emul, oracle, dim, bits, vec, pids, nbr, want_blocks=True, global_tile=True)
    assert np.array_equal(codes, want_c) and np.array_equal(aux.view(np.uint32), want_a.view(np.uint32))
    lay = co.nb_layout(max(16, 1 << (dim - 1).bit_length()), bits)
    want_b = common.blocks_from_codes(dim, bits, want_c, want_a, nbr, len(vec))
    assert np.array_equal(blocks[:, :lay["count"] + 4], want_b[:, :lay["count"] + 4])


@pytest.mark.parametrize("dim,bits,rows", [(960, 2, 32), (960, 4, 32), (1500, 1, 16), (1500, 4, 16)])
def test_emulated_kernel_large_padded_dims(emul, oracle, dim, bits, rows):
    """D = 1024 fills one CTA with one warp's tile; D = 2048 takes the 32 neighbours in two passes of 16 lanes."""
    vec, pids, nbr = common.neighbor_code_case(dim, 2, dim + bits)
    want_c, want_a = common.expected_neighbor_codes(oracle, dim, bits, vec, pids, nbr)
    codes, aux, used = _run(emul, oracle, dim, bits, vec, pids, nbr)
    assert used == rows
    assert np.array_equal(codes, want_c) and np.array_equal(aux.view(np.uint32), want_a.view(np.uint32))


def test_emulated_kernel_launch_shapes(emul, oracle):
    """Identity parent list, a row stride beyond dim, one warp per CTA and a parent count that leaves warps idle."""
    dim, bits = 96, 4
    vec, _, nbr = common.neighbor_code_case(dim, 11, 5)
    pids = np.arange(11, dtype=np.uint32)
    want_c, want_a = common.expected_neighbor_codes(oracle, dim, bits, vec, pids, nbr)
    for kw in ({}, {"max_warps": 1}, {"stride": 104}):
        codes, aux, _ = _run(emul, oracle, dim, bits, vec, None, nbr, **kw)
        assert np.array_equal(codes, want_c) and np.array_equal(aux.view(np.uint32), want_a.view(np.uint32)), kw


@pytest.mark.skipif(not co.have_ref(), reason="oracle/_ref (compiled reference) not present")
@pytest.mark.parametrize("bits", [1, 2, 4])
def test_emulated_kernel_equals_the_compiled_reference(emul, oracle, bits):
    vec, pids, nbr = common.neighbor_code_case(128, 16, 40 + bits)
    want_c, want_a = common.expected_neighbor_codes(oracle, 128, bits, vec, pids, nbr, ref=True)
    codes, aux, _ = _run(emul, oracle, 128, bits, vec, pids, nbr)
    assert np.array_equal(codes, want_c) and np.array_equal(aux.view(np.uint32), want_a.view(np.uint32))


@pytest.mark.parametrize("bits", [1, 2, 4])
@pytest.mark.parametrize("dim", [20, 128, 1500])
def test_emulated_block_output_is_the_codes_in_the_reference_layout(emul, oracle, dim, bits):
    vec, pids, nbr = common.neighbor_code_case(dim, 3, 9 * dim + bits)
    nbr[1, 20:] = 0xFFFFFFFF                            # a short list: count = one past the last occupied slot
    nbr[1, 19] = 5
    codes, aux, _, blocks = _run(emul, oracle, dim, bits, vec, pids, nbr, want_blocks=True)
    want = common.blocks_from_codes(dim, bits, codes, aux, nbr, len(vec))
    lay = co.nb_layout(max(16, 1 << (dim - 1).bit_length()), bits)
    assert np.array_equal(blocks[:, :lay["count"] + 4], want[:, :lay["count"] + 4])
    assert (blocks[:, lay["count"] + 4:] == 0xCC).all()   # the struct's tail padding is left alone
    assert blocks[1, lay["count"]:lay["count"] + 4].view(np.uint32)[0] == 20


@pytest.mark.parametrize("bits", [1, 2, 4])
def test_emulated_kernel_regenerates_the_blocks_of_a_reference_built_index(emul, oracle, bits):
    """Every neighbour block of a real index (built and saved by the unmodified reference), from nothing but its raw
    vectors and neighbour ids: what prune_and_write stored, bit for bit, up to each block's count."""
    sf = co.SaveFile(common.GOLDEN / f"ref_n300_d24_b{bits}.bin")
    ids, cnt = sf.field("ids").copy(), sf.field("count")
    for p in range(sf.n):
        ids[p, cnt[p]:] = 0xFFFFFFFF                    # slots past count hold leftovers of earlier refinement rounds
    vec = np.ascontiguousarray(sf.raw[:, :sf.dim])
    _, _, _, blocks = _run(emul, oracle, sf.dim, bits, vec, None, ids, want_blocks=True)
    lay, nb = sf.lay, sf.nb_off
    ref = sf.search_data[:, nb:]
    for p in range(sf.n):
        c = int(cnt[p])
        got_pl = blocks[p, :4 * sf.D * bits].reshape(bits, sf.D // 8, 32)
        ref_pl = ref[p, :4 * sf.D * bits].reshape(bits, sf.D // 8, 32)
        assert np.array_equal(got_pl[:, :, :c], ref_pl[:, :, :c]), p
        for name, width in (("nop", 4), ("ip_qo", 4), ("ip_cp", 4), ("pop", 2), ("wpop", 2), ("ids", 4)):
            if lay[name] is None:
                continue
            o = lay[name]
            assert np.array_equal(blocks[p, o:o + width * c], ref[p, o:o + width * c]), (p, name)
        assert blocks[p, lay["count"]:lay["count"] + 4].view(np.uint32)[0] == c
