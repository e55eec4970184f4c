"""Kernel-level hooks of the C ABI as torch-tensor functions (parity tests, micro-benchmarks).

Each function launches exactly one of the library's kernels on tensors that already live on the
index's GPU; torch is only the buffer allocator here.
"""
from __future__ import annotations

import torch

from . import _capi


def _dev(ix):
    return torch.device("cuda", ix._device)


def _stream(ix):
    return torch.cuda.current_stream(_dev(ix)).cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def prepare_queries(ix, queries: torch.Tensor, center: bool = False):
    """K1.  Returns dict(lut u8 [nq,D/4,16], coeffs f32 [nq,3], rotated f32 [nq,D], uplanes i32 [nq,4,W])."""
    info = ix.info()
    D, nq = info["D"], queries.shape[0]
    W = max(D, 128) // 32
    q = queries.to(_dev(ix), torch.float32).contiguous()
    out = {
        "lut": torch.empty((nq, D // 4, 16), dtype=torch.uint8, device=q.device),
        "coeffs": torch.empty((nq, 3), dtype=torch.float32, device=q.device),
        "rotated": torch.empty((nq, D), dtype=torch.float32, device=q.device),
        "uplanes": torch.empty((nq, 4, W), dtype=torch.int32, device=q.device),
    }
    _capi.check(ix.handle, ix._lib.cphnsw_b200_prepare_queries(
        ix.handle, q.data_ptr(), nq, int(center), out["lut"].data_ptr(), out["coeffs"].data_ptr(),
        out["rotated"].data_ptr(), out["uplanes"].data_ptr(), _stream(ix)))
    return out


def fastscan_blocks(ix, uplanes, coeffs, dqp, vertex_ids=None, first_vertex=0, nblocks=None,
                    query_of_block=None, slack_level=None, want=("nbit", "msb", "msb2", "est", "lower", "msb_lower"), out=None):
    """K2 over the neighbour blocks of `vertex_ids` (or a contiguous range).  Outputs are [nblocks, 32]; `out` may
    carry pre-allocated output tensors (same names) so that a timed call launches nothing but the kernel."""
    dev = _dev(ix)
    if vertex_ids is not None:
        vertex_ids = vertex_ids.to(dev, torch.int32).contiguous()
        nblocks = vertex_ids.numel()
    dqp = dqp.to(dev, torch.float32).contiguous()
    assert dqp.numel() == nblocks
    if query_of_block is not None:
        query_of_block = query_of_block.to(dev, torch.int32).contiguous()
    if slack_level is not None:
        slack_level = slack_level.to(dev, torch.int32).contiguous()
    given = out or {}
    out = {}
    for name in ("nbit", "msb", "msb2"):
        out[name] = (given.get(name) if name in given else torch.empty((nblocks, 32), dtype=torch.int32, device=dev)) if name in want else None
    for name in ("est", "lower", "msb_lower"):
        out[name] = (given.get(name) if name in given else torch.empty((nblocks, 32), dtype=torch.float32, device=dev)) if name in want else None
    _capi.check(ix.handle, ix._lib.cphnsw_b200_fastscan_blocks(
        ix.handle, uplanes.data_ptr(), coeffs.data_ptr(), uplanes.shape[0], _ptr(query_of_block), _ptr(vertex_ids),
        first_vertex, nblocks, dqp.data_ptr(), _ptr(slack_level), _ptr(out["nbit"]), _ptr(out["msb"]),
        _ptr(out["msb2"]), _ptr(out["est"]), _ptr(out["lower"]), _ptr(out["msb_lower"]), _stream(ix)))
    return {k: v for k, v in out.items() if v is not None}


def exact_l2(ix, queries: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    """K4 primitive: exact distances of ids [nq, m] to queries [nq, dim]."""
    dev = _dev(ix)
    q = queries.to(dev, torch.float32).contiguous()
    ids = ids.to(dev, torch.int32).contiguous()
    out = torch.empty(ids.shape, dtype=torch.float32, device=dev)
    _capi.check(ix.handle, ix._lib.cphnsw_b200_exact_l2(
        ix.handle, q.data_ptr(), q.shape[0], ids.data_ptr(), ids.shape[1], out.data_ptr(), _stream(ix)))
    return out


def greedy_descent(ix, queries: torch.Tensor) -> torch.Tensor:
    """K3 prologue: layer-0 entry point per query."""
    dev = _dev(ix)
    q = queries.to(dev, torch.float32).contiguous()
    out = torch.empty((q.shape[0],), dtype=torch.int32, device=dev)
    _capi.check(ix.handle, ix._lib.cphnsw_b200_greedy_descent(ix.handle, q.data_ptr(), q.shape[0], out.data_ptr(), _stream(ix)))
    return out


def exhaustive_estimates(ix, queries: torch.Tensor, id_begin: int = 0, id_end: int | None = None):
    dev = _dev(ix)
    q = queries.to(dev, torch.float32).contiguous()
    id_end = ix.size if id_end is None else id_end
    m = id_end - id_begin
    sums = torch.empty((q.shape[0], m), dtype=torch.int32, device=dev)
    est = torch.empty((q.shape[0], m), dtype=torch.float32, device=dev)
    _capi.check(ix.handle, ix._lib.cphnsw_b200_exhaustive_estimates(
        ix.handle, q.data_ptr(), q.shape[0], id_begin, id_end, sums.data_ptr(), est.data_ptr(), _stream(ix)))
    return sums, est


def exhaustive_search(ix, queries: torch.Tensor, k: int, kprime: int, id_begin: int = 0, id_end: int | None = None):
    dev = _dev(ix)
    q = queries.to(dev, torch.float32).contiguous()
    id_end = ix.size if id_end is None else id_end
    ids = torch.empty((q.shape[0], k), dtype=torch.int64, device=dev)
    dists = torch.empty((q.shape[0], k), dtype=torch.float32, device=dev)
    _capi.check(ix.handle, ix._lib.cphnsw_b200_exhaustive_search(
        ix.handle, q.data_ptr(), q.shape[0], k, kprime, id_begin, id_end, ids.data_ptr(), dists.data_ptr(), _stream(ix)))
    return ids, dists


def neighbor_block_bytes(D: int, bits: int) -> int:
    """sizeof(FastScanNeighborBlock<D>) / sizeof(NbitFastScanNeighborBlock<D, 32, bits>)."""
    return -(-(4 * D * bits + 384 + 64 * (2 if bits > 1 else 1) + 128 + 4) // 64) * 64


def neighbor_codes(ix, vectors: torch.Tensor, nbr_ids: torch.Tensor, parent_ids: torch.Tensor | None = None,
                   rotation_seed: int = 42, blocks: bool = False):
    """N3 (build side).  vectors f32 [n, dim], nbr_ids i32 [n_parents, 32] (-1 = empty slot), parent_ids i32 [n_parents]
    (None = 0, 1, ...) -> (codes u8 [n_parents, 32, bits, D/8], aux f32 [n_parents, 32, 3] = nop, ip_qo, ip_cp) for the
    dim and bits `ix` was created with; blocks=True: also the neighbour blocks in the reference's layout, u8
    [n_parents, neighbor_block_bytes(D, bits)] (zero-filled first).  The index needs no data on the device."""
    dev = _dev(ix)
    dim, bits = ix.dim, ix._bits
    D = max(16, 1 << (dim - 1).bit_length())
    v = vectors.to(dev, torch.float32).contiguous()
    assert v.dim() == 2 and v.shape[1] == dim
    nb = nbr_ids.to(dev, torch.int32).contiguous()
    assert nb.dim() == 2 and nb.shape[1] == 32
    pid = None if parent_ids is None else parent_ids.to(dev, torch.int32).contiguous()
    assert pid is None or pid.numel() == nb.shape[0]
    codes = torch.empty((nb.shape[0], 32, bits, D // 8), dtype=torch.uint8, device=dev)
    aux = torch.empty((nb.shape[0], 32, 3), dtype=torch.float32, device=dev)
    blk = torch.zeros((nb.shape[0], neighbor_block_bytes(D, bits)), dtype=torch.uint8, device=dev) if blocks else None
    _capi.check(ix.handle, ix._lib.cphnsw_b200_neighbor_codes(
        ix.handle, dim, bits, rotation_seed, v.data_ptr(), v.shape[1], v.shape[0], _ptr(pid), nb.data_ptr(), nb.shape[0],
        codes.data_ptr(), aux.data_ptr(), _ptr(blk), 0 if blk is None else blk.shape[1], _stream(ix)))
    return (codes, aux, blk) if blocks else (codes, aux)


def exhaustive_candidates(ix, queries: torch.Tensor, kprime: int, id_begin: int, id_end: int, id_offset: int = 0,
                          prior_keys: torch.Tensor | None = None, tau_in: torch.Tensor | None = None, want_dists: bool = False):
    """One piece of a scan done in pieces (cphnsw_b200_exhaustive_candidates): the kprime smallest (estimate, id) keys of
    `prior_keys` U [id_begin, id_end) as an int64 tensor [nq, kprime] holding the uint64 bit patterns (estimate bits << 32 |
    id; -1 = padding), their exact distances [nq, kprime] when want_dists (the last piece: ids are then shifted by
    id_offset), and tau [nq] = the estimate of the kprime-th key (FLT_MAX while fewer).  tau_in [nq]: bounds from elsewhere."""
    dev = _dev(ix)
    q = queries.to(dev, torch.float32).contiguous()
    nq = q.shape[0]
    keys = torch.empty((nq, kprime), dtype=torch.int64, device=dev)
    dists = torch.empty((nq, kprime), dtype=torch.float32, device=dev) if want_dists else None
    tau = torch.empty((nq,), dtype=torch.float32, device=dev)
    pk = None if prior_keys is None else prior_keys.to(dev, torch.int64).contiguous()
    ti = None if tau_in is None else tau_in.to(dev, torch.float32).contiguous()
    assert pk is None or tuple(pk.shape) == (nq, kprime)
    assert ti is None or tuple(ti.shape) == (nq,)
    _capi.check(ix.handle, ix._lib.cphnsw_b200_exhaustive_candidates(
        ix.handle, q.data_ptr(), nq, kprime, id_begin, id_end, id_offset, _ptr(pk), _ptr(ti), keys.data_ptr(), _ptr(dists),
        tau.data_ptr(), _stream(ix)))
    return keys, dists, tau


def merge_candidates(ix, keys: torch.Tensor, dists: torch.Tensor | None, k: int):
    """cphnsw_b200_merge_candidates: keys int64 [lists, nq, kprime] (+ dists float32 alongside) -> (ids int64 [nq, k], dists
    float32 [nq, k], tau float32 [nq]); with dists None or k == 0 only tau (ids, dists are None)."""
    dev = _dev(ix)
    kk = keys.to(dev, torch.int64).contiguous()
    lists, nq, kp = kk.shape
    dd = None if dists is None else dists.to(dev, torch.float32).contiguous()
    tau = torch.empty((nq,), dtype=torch.float32, device=dev)
    want = dd is not None and k > 0
    ids = torch.empty((nq, k), dtype=torch.int64, device=dev) if want else None
    out = torch.empty((nq, k), dtype=torch.float32, device=dev) if want else None
    _capi.check(ix.handle, ix._lib.cphnsw_b200_merge_candidates(
        ix.handle, kk.data_ptr(), _ptr(dd), lists, nq, kp, k if want else 0, _ptr(ids), _ptr(out), tau.data_ptr(), _stream(ix)))
    return ids, out, tau


def calibration_samples(ix, queries: torch.Tensor, start_ids: torch.Tensor):
    """N4: the sample loop of Index::calibrate_estimator (cphnsw_b200_calibration_samples).  queries f32 [ns, dim], start_ids
    i32 [ns] -> dict of parent i32 [ns], nn_dist_sq / dist_qp_sq f32 [ns], nop / ip_corrected / ip_qo_denom / true_ip f32
    [ns, 32], neighbor i32 [ns, 32] (-1 past the parent's last neighbour)."""
    dev = _dev(ix)
    q = queries.to(dev, torch.float32).contiguous()
    st = start_ids.to(dev, torch.int32).contiguous()
    ns = q.shape[0]
    assert tuple(st.shape) == (ns,) and q.shape[1] == ix.dim
    o = {"parent": torch.empty((ns,), dtype=torch.int32, device=dev), "nn_dist_sq": torch.empty((ns,), dtype=torch.float32, device=dev),
         "dist_qp_sq": torch.empty((ns,), dtype=torch.float32, device=dev)}
    for name in ("nop", "ip_corrected", "ip_qo_denom", "true_ip"):
        o[name] = torch.empty((ns, 32), dtype=torch.float32, device=dev)
    o["neighbor"] = torch.empty((ns, 32), dtype=torch.int32, device=dev)
    _capi.check(ix.handle, ix._lib.cphnsw_b200_calibration_samples(
        ix.handle, q.data_ptr(), st.data_ptr(), ns, o["parent"].data_ptr(), o["nn_dist_sq"].data_ptr(), o["dist_qp_sq"].data_ptr(),
        o["nop"].data_ptr(), o["ip_corrected"].data_ptr(), o["ip_qo_denom"].data_ptr(), o["true_ip"].data_ptr(), o["neighbor"].data_ptr(),
        _stream(ix)))
    return o


def upload_arrays(ix, *, D, bits, dim, search_data, raw, norm_sq, calibration, centroid=None, max_level=0,
                  entry_point=0, graph_entry_point=0, rotation_seed=42, layers=()):
    """cphnsw_b200_upload from numpy arrays laid out like the reference's in-memory index.

    search_data: uint8 [n, rec_size] VertexSearchData records; raw: float32 [n, D]; calibration: the
    248-byte CalibrationSnapshot; layers: sequence of (nodes, offsets, neighbours) uint32 arrays.
    """
    import ctypes as C

    import numpy as np

    n = search_data.shape[0]
    sd = np.ascontiguousarray(search_data, np.uint8)
    raw = np.ascontiguousarray(raw, np.float32)
    ns = np.ascontiguousarray(norm_sq, np.float32)
    cal = np.ascontiguousarray(np.frombuffer(bytes(calibration), np.uint8))
    cen = np.ascontiguousarray(np.zeros(dim, np.float32) if centroid is None else centroid, np.float32)
    words = (D + 63) // 64
    storage = -(-(8 * words * bits) // 64) * 64
    nb_off = -(-(storage + 8) // 64) * 64
    h = _capi.HostIndex()
    h.D, h.bits, h.dim, h.n = D, bits, dim, n
    h.search_data, h.rec_size, h.nb_off = sd.ctypes.data, sd.shape[1], nb_off
    h.raw, h.norm_sq, h.centroid, h.calibration = raw.ctypes.data, ns.ctypes.data, cen.ctypes.data, cal.ctypes.data
    h.max_level, h.entry_point, h.graph_entry_point, h.rotation_seed = max_level, entry_point, graph_entry_point, rotation_seed
    nl = len(layers)
    h.n_layers = nl
    u32p = C.POINTER(C.c_uint32)
    keep = []
    arrs = [(u32p * max(nl, 1))() for _ in range(3)]
    sizes = np.zeros(max(nl, 1), np.uint32)
    for i, (nodes, offs, nbrs) in enumerate(layers):
        nodes = np.ascontiguousarray(nodes, np.uint32)
        offs = np.ascontiguousarray(offs, np.uint32)
        nbrs = np.ascontiguousarray(nbrs if len(nbrs) else np.zeros(1, np.uint32), np.uint32)
        keep += [nodes, offs, nbrs]
        for a, x in zip(arrs, (nodes, offs, nbrs)):
            a[i] = x.ctypes.data_as(u32p)
        sizes[i] = nodes.size
    h.layer_nodes, h.layer_offs, h.layer_nbrs = (C.cast(a, C.POINTER(u32p)) for a in arrs)
    h.layer_sizes = sizes.ctypes.data_as(u32p)
    _capi.check(ix.handle, ix._lib.cphnsw_b200_upload(ix.handle, C.byref(h)))
    info = _capi.Info()
    _capi.check(ix.handle, ix._lib.cphnsw_b200_get_info(ix.handle, C.byref(info)))
    ix._info = info
    ix._finalized = True
    ix._source_path = None
