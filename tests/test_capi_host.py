"""CPU: the C-ABI library loads and exports every symbol include/cphnsw_b200.h declares; the host-side
mirror of the reference interface validates arguments like the reference; without a GPU every
entry point fails loudly (no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import common

ROOT = Path(__file__).resolve().parents[1]


def _declared():
    text = (ROOT / "include" / "cphnsw_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cphnsw_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from cphnsw_b200 import _capi

    lib = _capi.lib()
    names = _declared()
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/cphnsw_b200.h but not exported"
        assert n in _capi.SYMBOLS, f"{n} has no ctypes signature"
    assert set(_capi.SYMBOLS) <= set(names)


def test_header_cites_the_reference_interface():
    text = (ROOT / "include" / "cphnsw_b200.h").read_text()
    for cite in ("src/bindings.cpp", "api/hnsw_index.hpp", "distance/fastscan_kernel.hpp", "search/rabitq_search.hpp",
                 "encoder/rabitq_encoder.hpp", "core/memory.hpp"):
        assert cite in text


def test_the_product_never_imports_the_oracle():
    pkg = ROOT / "rabitq-ann-search_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.h")):
        src = p.read_text()
        # comments may cite the probe that pinned a float sequence; code may not import / include / call it
        assert "cphnsw_oracle" not in src and "cpo_" not in src and "libcphnsw_refshim" not in src, p
        assert not re.search(r'#include\s*[<"][^>"]*oracle', src) and not re.search(r"^\s*(from|import)\s+oracle", src, re.M), p


@pytest.mark.skipif(common.has_cuda(), reason="this is the no-GPU behaviour")
def test_no_gpu_means_loud_failure_not_fallback():
    import cphnsw_b200
    from cphnsw_b200 import _capi

    lib = _capi.lib()
    h = C.c_void_p()
    rc = lib.cphnsw_b200_create(0, C.byref(h))
    assert rc == _capi.ECUDA and not h.value
    assert b"no CPU path" in lib.cphnsw_b200_last_error(None)
    with pytest.raises(RuntimeError, match="no CPU path"):
        cphnsw_b200.CPIndex(128, 4)


def test_a_missing_library_is_an_import_error_not_a_fallback(tmp_path):
    """CPHNSW_B200_LIB names another build of the native library (A/B runs of kernel variants); a path with nothing behind it
    must fail as loudly as a missing default build."""
    import subprocess
    import sys

    code = ("import sys; sys.path.insert(0, %r)\n"
            "from cphnsw_b200 import _capi\n"
            "try:\n    _capi.lib()\nexcept ImportError as e:\n    print('IMPORT_ERROR', 'no CPU fallback' in str(e))\n") % str(ROOT / "rabitq-ann-search_b200")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True,
                       env=dict(__import__("os").environ, CPHNSW_B200_LIB=str(tmp_path / "libnothing.so")))
    assert "IMPORT_ERROR True" in r.stdout, (r.stdout, r.stderr[-500:])


def test_constructor_validation_matches_the_reference_factory():
    import cphnsw_b200

    # these are raised before any device is touched (src/bindings.cpp:77-113)
    with pytest.raises(ValueError, match=r"Unsupported bits=3\. Supported: 1, 2, 4\."):
        cphnsw_b200.CPIndex(128, 3)
    with pytest.raises(ValueError, match=r"Unsupported dimension 5000 \(padded to 8192\)"):
        cphnsw_b200.CPIndex(5000, 1)
    with pytest.raises(ValueError, match="Unsupported dimension"):
        cphnsw_b200.CPIndex(0, 1)


def test_struct_layouts_match_the_header(tmp_path):
    """The header is valid C, and the ctypes mirrors have the sizes / offsets the C compiler gives."""
    import subprocess

    from cphnsw_b200 import _capi

    src = tmp_path / "sz.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "cphnsw_b200.h"\n'
        'int main(void){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(cphnsw_b200_info), sizeof(cphnsw_b200_stats), '
        'sizeof(cphnsw_b200_host_index), offsetof(cphnsw_b200_host_index, layer_sizes), '
        'offsetof(cphnsw_b200_info, slack_levels), offsetof(cphnsw_b200_host_index, rotation_seed));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    got = list(map(int, subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()))
    assert got == [C.sizeof(_capi.Info), C.sizeof(_capi.Stats), C.sizeof(_capi.HostIndex), _capi.HostIndex.layer_sizes.offset,
                   _capi.Info.slack_levels.offset, _capi.HostIndex.rotation_seed.offset]


def test_query_sharding_plan():
    from cphnsw_b200 import sharding

    for nq, world in ((10, 1), (10, 3), (7, 8), (0, 4), (10000, 8)):
        parts = [sharding.query_shard(nq, r, world) for r in range(world)]
        assert parts[0][0] == 0 and parts[-1][1] == nq
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        sizes = [e - b for b, e in parts]
        assert max(sizes) - min(sizes) <= 1
    for n, world in ((100, 4), (10_000_001, 8), (5, 8)):
        parts = [sharding.db_shard(n, r, world) for r in range(world)]
        assert parts[0][0] == 0 and parts[-1][1] == n and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))


def test_topk_merge_is_exact():
    from cphnsw_b200 import sharding

    rng = np.random.default_rng(0)
    nq, k, shards = 13, 10, 4
    d = rng.random((shards, nq, k)).astype(np.float32)
    d.sort(axis=2)
    d[1, :, 3] = d[0, :, 3]          # cross-shard ties: broken by id
    ids = rng.permutation(shards * nq * k).reshape(shards, nq, k).astype(np.int64)
    ids[2, 0, 7:] = -1
    d[2, 0, 7:] = np.finfo(np.float32).max
    mi, md = sharding.merge_topk(ids, d, k)
    for qi in range(nq):
        cand = [(float(d[s, qi, j]), int(ids[s, qi, j])) for s in range(shards) for j in range(k) if ids[s, qi, j] >= 0]
        cand.sort()
        assert [c[1] for c in cand[:k]] == mi[qi].tolist()
        assert np.array_equal(np.float32([c[0] for c in cand[:k]]), md[qi])


def test_internal_to_original_id_map_from_a_reference_save_file():
    """recover_id_map on the committed reference-written index files: the map is a permutation of the rows the index
    was built from (tests/golden/make_golden.py: co.synthetic(300, 24, seed=7)) and reproduces the stored vectors."""
    src = (ROOT / "rabitq-ann-search_b200" / "cphnsw_b200" / "index.py").read_text()
    ns = {"__name__": "idx_only"}
    exec(compile(src.replace("from . import _capi", "_capi = None"), "index.py", "exec"), ns)   # the map needs no native code
    rng = np.random.default_rng(7)
    base = rng.standard_normal((300, 24)).astype(np.float32)
    for bits in (1, 2, 4):
        path = str(ROOT / "tests" / "golden" / f"ref_n300_d24_b{bits}.bin")
        m = ns["recover_id_map"](path, base, 24)
        assert sorted(m.tolist()) == list(range(300))
        raw = np.memmap(path, np.float32, "r", 68 + 248 + 72 + 4 * 24 + 8 * 300, (300, 32))[:, :24]
        assert np.array_equal(base[m], raw)
    with pytest.raises(ValueError):
        ns["recover_id_map"](path, base[::-1] * 2, 24)
