"""Shared test helpers: synthetic indexes in the reference's record layout, oracle views, caches.

Everything here is test infrastructure; it may use oracle/ (the product never does).
"""
from __future__ import annotations

import os
import struct
import sys
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
for _p in (ROOT / "oracle", ROOT / "rabitq-ann-search_b200"):
    if str(_p) not in sys.path:
        sys.path.insert(0, str(_p))

import cphnsw_oracle as co  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"
CACHE = Path(os.environ.get("CPHNSW_B200_TEST_CACHE", "/tmp/cphnsw_b200_test_cache"))


def has_cuda() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        return False


def calibration_bytes(a=1.0, b=0.0, floor=0.3, slacks=(0.8, 0.85, 0.9), gamma=1.5, gamma_max=3.0, gamma_beta=0.5,
                      gamma_warmup=8) -> bytes:
    """A CalibrationSnapshot (api/hnsw_index.hpp:33-58; offsets SURVEY App. B), 248 bytes."""
    buf = bytearray(248)
    struct.pack_into("<3f", buf, 0, a, b, floor)
    struct.pack_into("<3f", buf, 80, gamma, gamma_max, gamma_beta)   # gamma_min := gamma
    struct.pack_into("<Q", buf, 96, gamma_warmup)
    struct.pack_into("<i", buf, 104, len(slacks))
    sl = list(slacks) + [0.0] * (32 - len(slacks))
    struct.pack_into("<32f", buf, 108, *sl)
    struct.pack_into("<i", buf, 236, len(slacks))
    struct.pack_into("<f", buf, 240, gamma)
    return bytes(buf)


def fabricate(n, dim, bits, seed=0, counts=(32,), degenerate=False, layers=0, **calib):
    """A random 'finalized index' in the reference's in-memory layout.

    Codes, aux values and the graph are random, so estimates are meaningless -- but the query path is
    a deterministic function of these bytes, which is all parity needs, and unlike real indexes this
    reaches every branch: partial blocks (count % 8 tails, count 0), ip_qo under the floor / zero,
    gamma termination, MSB-only skips, duplicate vectors.
    """
    rng = np.random.default_rng(seed)
    D = 1 << (dim - 1).bit_length()
    nb_off = co.code_bytes(D, bits)
    lay = co.nb_layout(D, bits)
    rec = nb_off + lay["size"]
    sd = np.zeros((n, rec), np.uint8)
    nsp = D // 8
    planes = rng.integers(0, 256, (n, bits, nsp, 32), dtype=np.uint8)
    if D > dim:   # padded dims carry no code bits in a real index; keep some anyway in half the blocks
        pass
    sd[:, nb_off:nb_off + 4 * D * bits] = planes.reshape(n, -1)
    cnt = rng.choice(np.asarray(counts, np.uint32), n)
    ids = np.empty((n, 32), np.uint32)
    for v in range(n):
        ids[v] = rng.choice(n, 32, replace=n < 32)
    slot = np.arange(32)[None, :]
    ids[slot >= cnt[:, None]] = 0xFFFFFFFF
    nop = rng.uniform(0.4, 2.5, (n, 32)).astype(np.float32)
    ip_qo = rng.uniform(0.2, 1.1, (n, 32)).astype(np.float32)
    ip_cp = rng.normal(0, 0.05, (n, 32)).astype(np.float32)
    if degenerate:
        m = rng.random((n, 32))
        ip_qo[m < 0.1] = 0.0
        ip_qo[(m >= 0.1) & (m < 0.15)] = 1e-12
        nop[m > 0.97] = 0.0
    bitcnt = np.unpackbits(planes, axis=2).reshape(n, bits, nsp, 8, 32).sum(axis=(2, 3)).astype(np.uint32)  # [n,bits,32]
    pop = bitcnt[:, 0, :].astype(np.uint16)
    wpop = sum(bitcnt[:, b, :] << (bits - 1 - b) for b in range(bits)).astype(np.uint16)

    def put(name, arr):
        o = nb_off + lay[name]
        b = np.ascontiguousarray(arr).view(np.uint8).reshape(n, -1)
        sd[:, o:o + b.shape[1]] = b

    put("nop", nop); put("ip_qo", ip_qo); put("ip_cp", ip_cp); put("pop", pop)
    if bits > 1:
        put("wpop", wpop)
    put("ids", ids)
    put("count", cnt.astype(np.uint32).reshape(n, 1))
    # per-vertex code (bits == 1: signs, nop, ip_qo) for the exhaustive scan
    if bits == 1:
        storage = -(-(8 * ((D + 63) // 64)) // 64) * 64
        signs = rng.integers(0, 256, (n, D // 8), dtype=np.uint8)
        sd[:, :D // 8] = signs
        vn = rng.uniform(0.5, 3.0, n).astype(np.float32)
        vq = rng.uniform(0.3, 1.0, n).astype(np.float32)
        sd[:, storage:storage + 4] = vn.view(np.uint8).reshape(n, 4)
        sd[:, storage + 4:storage + 8] = vq.view(np.uint8).reshape(n, 4)
    raw = np.zeros((n, D), np.float32)
    raw[:, :dim] = rng.standard_normal((n, dim)).astype(np.float32)
    if degenerate and n > 8:
        raw[5] = raw[3]   # duplicate vectors: exact ties between distinct ids
        raw[6] = raw[3]
    norm_sq = np.zeros(n, np.float32)
    for i in range(dim):   # sequential f32 sum like rabitq_graph.hpp:84-87
        norm_sq += raw[:, i] * raw[:, i]
    lay_list = []
    level_nodes = np.arange(n, dtype=np.uint32)
    for _L in range(layers):
        keep = max(2, len(level_nodes) // 6)
        level_nodes = np.sort(rng.choice(level_nodes, keep, replace=False)).astype(np.uint32)
        offs = np.zeros(keep + 1, np.uint32)
        nbrs = []
        for i, _node in enumerate(level_nodes):
            deg = int(rng.integers(0, min(9, keep)))
            nb = rng.choice(level_nodes, deg, replace=False)
            nbrs.append(nb.astype(np.uint32))
            offs[i + 1] = offs[i] + deg
        lay_list.append((level_nodes, offs, np.concatenate(nbrs) if nbrs else np.zeros(0, np.uint32)))
    entry = int(level_nodes[0]) if layers else int(rng.integers(0, n))
    cal = calibration_bytes(**calib)
    centroid = rng.normal(0, 0.1, dim).astype(np.float32)
    return types.SimpleNamespace(
        D=D, B=bits, dim=dim, n=n, R=32, search_data=sd, raw=raw, norm_sq=norm_sq, rec_size=rec, nb_off=nb_off,
        lay=lay, calibration=cal, centroid=centroid, max_level=layers, entry_point=entry, rotation_seed=42,
        layers=lay_list, **_calib_fields(cal))


def _calib_fields(cal: bytes) -> dict:
    f32 = lambda o: float(np.frombuffer(cal, np.float32, 1, o)[0])  # noqa: E731
    return dict(affine_a=f32(0), affine_b=f32(4), ip_qo_floor=f32(8), gamma_min=f32(80), gamma_max=f32(84),
                gamma_beta=f32(88), gamma_warmup=int.from_bytes(cal[96:104], "little"),
                slack_levels=np.frombuffer(cal, np.float32, 32, 108).copy(),
                num_slack_levels=int.from_bytes(cal[236:240], "little", signed=True), search_gamma=f32(240))


def write_save_file(fab, path):
    """Serialise a fabricated index as a save file v2 (api/hnsw_index.hpp:217-303)."""
    with open(path, "wb") as f:
        f.write(struct.pack("<QIIIIIQiIffdQ", co.MAGIC, 2, fab.D, 32, fab.B, fab.dim, fab.n, fab.max_level,
                            fab.entry_point, 0.0, 1.0, 0.5, fab.rotation_seed))
        f.write(fab.calibration)
        f.write(bytes(72))
        f.write(np.ascontiguousarray(fab.centroid, np.float32).tobytes())
        f.write(np.zeros(fab.n, np.int32).tobytes())
        f.write(np.ascontiguousarray(fab.norm_sq, np.float32).tobytes())
        f.write(np.ascontiguousarray(fab.raw, np.float32).tobytes())
        f.write(np.ascontiguousarray(fab.search_data).tobytes())
        f.write(struct.pack("<I", len(fab.layers)))
        for nodes, offs, nbrs in fab.layers:
            f.write(struct.pack("<I", len(nodes)))
            for i, node in enumerate(nodes):
                nb = nbrs[offs[i]:offs[i + 1]]
                f.write(struct.pack("<II", int(node), len(nb)))
                f.write(np.ascontiguousarray(nb, np.uint32).tobytes())
    return path


def gpu_index_from(fab_or_path, dim=None, bits=None):
    """A cphnsw_b200.CPIndex holding `fab_or_path` (a fabricate() result or a save-file path)."""
    import cphnsw_b200
    from cphnsw_b200 import hooks

    if isinstance(fab_or_path, (str, Path)):
        sf = co.SaveFile(fab_or_path)
        ix = cphnsw_b200.CPIndex(sf.dim, sf.B)
        ix.load(str(fab_or_path))
        ix.set_option("collect_stats", 1)
        return ix
    fab = fab_or_path
    ix = cphnsw_b200.CPIndex(fab.dim, fab.B)
    hooks.upload_arrays(ix, D=fab.D, bits=fab.B, dim=fab.dim, search_data=fab.search_data, raw=fab.raw,
                        norm_sq=fab.norm_sq, calibration=fab.calibration, centroid=fab.centroid,
                        max_level=fab.max_level, entry_point=fab.entry_point, graph_entry_point=fab.entry_point,
                        rotation_seed=fab.rotation_seed, layers=fab.layers)
    ix.set_option("collect_stats", 1)
    return ix


def reference_index_file(n, dim, bits, clusters=0, seed=1234, threads=8):
    """Build (once per cache) a real index with the unmodified reference from oracle/_ref."""
    CACHE.mkdir(parents=True, exist_ok=True)
    path = CACHE / f"ref_{n}_{dim}_{bits}_{clusters}_{seed}.bin"
    if not path.exists():
        base = co.synthetic(n, dim, seed, clusters)
        tmp = str(path) + f".tmp{os.getpid()}"
        co.build_reference_index(base, bits, tmp, threads=threads)
        os.replace(tmp, path)
    return path


def queries_for(dim, nq, seed=99, clusters=0, base_seed=1234):
    if clusters:
        rng = np.random.default_rng(seed)
        centers = np.random.default_rng(base_seed).standard_normal((clusters, dim)).astype(np.float32) * np.float32(4.0)
        return (centers[rng.integers(0, clusters, nq)] + rng.standard_normal((nq, dim)).astype(np.float32)).astype(np.float32)
    return np.random.default_rng(seed).standard_normal((nq, dim)).astype(np.float32)


def sorted_rows(ids, dists):
    """Rows as sorted (dist, id) pairs: the reference's tie order between equal distances is unspecified."""
    order = np.lexsort((ids, dists), axis=1)
    return np.take_along_axis(ids, order, 1), np.take_along_axis(dists, order, 1)


# ---- N3 (build side): neighbour codes ------------------------------------------------------------
def neighbor_code_case(dim, n_parents, seed, holes=True):
    """vectors f32 [n, dim] (the second half are near-duplicates of the first: short offsets), parent ids u32
    [n_parents], neighbour ids u32 [n_parents, 32] with empty slots (0xFFFFFFFF) and one neighbour equal to its
    parent (nop == 0, the norm_epsilon branch)."""
    rng = np.random.default_rng(seed)
    n = n_parents * 8 + 40
    vec = rng.standard_normal((n, dim)).astype(np.float32)
    vec[n // 2:] = vec[:n - n // 2] + 0.05 * rng.standard_normal((n - n // 2, dim)).astype(np.float32)
    pids = rng.choice(n, n_parents, replace=False).astype(np.uint32)
    nbr = rng.integers(0, n, (n_parents, 32)).astype(np.uint32)
    if holes:
        nbr[rng.random((n_parents, 32)) < 0.15] = 0xFFFFFFFF
        nbr[0, 3] = pids[0]
    return vec, pids, nbr


def expected_neighbor_codes(oracle, dim, bits, vec, pids, nbr, ref=False):
    """What the oracle (ref=True: the compiled reference) writes for the case: codes u8 [np, 32, bits, D/8], aux f32 [np, 32, 3]."""
    D = max(16, 1 << (dim - 1).bit_length())
    codes = np.zeros((len(pids), 32, bits, D // 8), np.uint8)
    aux = np.zeros((len(pids), 32, 3), np.float32)
    for p in range(len(pids)):
        ok = nbr[p] < len(vec)
        nb = np.zeros((32, dim), np.float32)
        nb[ok] = vec[nbr[p][ok]]
        c, a = oracle.neighbor_aux(dim, bits, vec[pids[p]], nb, ref=ref)
        codes[p][ok] = c[ok]
        aux[p][ok] = a[ok]
    return codes, aux


def blocks_from_codes(dim, bits, codes, aux, nbr, n_vectors):
    """The reference's neighbour blocks (fastscan_layout.hpp:51-92, 114-155) holding `codes`/`aux` of one call: packed
    planes (byte sp of slot v at [plane][sp][v]), nop, ip_qo, ip_cp, popcounts, weighted popcounts, ids, count."""
    D = max(16, 1 << (dim - 1).bit_length())
    lay = co.nb_layout(D, bits)
    out = np.zeros((len(nbr), lay["size"]), np.uint8)
    pc = np.unpackbits(codes, axis=3).sum(3).astype(np.uint32)          # [np, 32, bits]
    for p in range(len(nbr)):
        out[p, :4 * D * bits] = np.transpose(codes[p], (1, 2, 0)).reshape(-1)
        for k, name in enumerate(("nop", "ip_qo", "ip_cp")):
            out[p, lay[name]:lay[name] + 128] = np.ascontiguousarray(aux[p, :, k]).view(np.uint8)
        out[p, lay["pop"]:lay["pop"] + 64] = pc[p, :, 0].astype(np.uint16).view(np.uint8)
        if bits > 1:
            w = sum(pc[p, :, b] << (bits - 1 - b) for b in range(bits))
            out[p, lay["wpop"]:lay["wpop"] + 64] = w.astype(np.uint16).view(np.uint8)
        ok = nbr[p] < n_vectors
        ids = np.where(ok, nbr[p], 0xFFFFFFFF).astype(np.uint32)
        out[p, lay["ids"]:lay["ids"] + 128] = ids.view(np.uint8)
        cnt = int(np.nonzero(ok)[0].max()) + 1 if ok.any() else 0
        out[p, lay["count"]:lay["count"] + 4] = np.array([cnt], np.uint32).view(np.uint8)
    return out
