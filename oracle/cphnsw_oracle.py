"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the parity oracle.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  It wraps

* ``libcphnsw_oracle.so``  -- our plain-C restatement of the hot path (oracle/cphnsw_oracle.c),
* ``_ref/libcphnsw_refshim.so`` and ``_ref/cphnsw/_core*.so`` -- the UNMODIFIED reference,
  compiled by oracle/Makefile (present only where it was built; they travel to the GPU box as
  prebuilt files),

and holds an independent numpy parser of the reference's save-file v2
(api/hnsw_index.hpp:217-303) used to hand a finalized reference index to
the C restatement.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_DIR = HERE / "_ref"
MAGIC = 0x57534E48504300

c_f32p = C.POINTER(C.c_float)
c_u8p = C.POINTER(C.c_uint8)
c_u16p = C.POINTER(C.c_uint16)
c_u32p = C.POINTER(C.c_uint32)
c_i64p = C.POINTER(C.c_int64)


def _ptr(a: np.ndarray, t):
    return a.ctypes.data_as(t)


def build_port(force: bool = False) -> Path:
    so = HERE / "libcphnsw_oracle.so"
    src = HERE / "cphnsw_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(HERE), "port"], check=True, capture_output=True)
    return so


def build_ref() -> bool:
    """Compile the unmodified reference into oracle/_ref (only where /root/reference exists)."""
    if not Path("/root/reference/src/bindings.cpp").exists():
        return have_ref()
    subprocess.run(["make", "-C", str(HERE), "ref"], check=True, capture_output=True)
    return have_ref()


def have_ref() -> bool:
    return (REF_DIR / "libcphnsw_refshim.so").exists() and any((REF_DIR / "cphnsw").glob("_core*.so"))


def ref_module():
    """The unmodified reference Python package (cphnsw.CPIndex), from oracle/_ref."""
    if str(REF_DIR) not in sys.path:
        sys.path.insert(0, str(REF_DIR))
    import cphnsw  # noqa: PLC0415

    return cphnsw


class CpoStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "pops", "expansions", "exact_calls", "beam_pushes", "max_beam", "nn_pushes",
        "lb_skips", "gamma_terms", "msb_skipped", "estimated", "descent_dists")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class CpoIndex(C.Structure):
    _fields_ = [
        ("D", C.c_uint32), ("B", C.c_uint32), ("dim", C.c_uint32), ("n", C.c_uint64),
        ("search_data", c_u8p), ("rec_size", C.c_uint64), ("nb_off", C.c_uint32),
        ("raw", c_f32p), ("norm_sq", c_f32p),
        ("affine_a", C.c_float), ("affine_b", C.c_float), ("ip_qo_floor", C.c_float),
        ("slack_levels", C.c_float * 32), ("num_slack_levels", C.c_int32),
        ("search_gamma", C.c_float), ("gamma_max", C.c_float), ("gamma_beta", C.c_float),
        ("gamma_warmup", C.c_uint64),
        ("max_level", C.c_int32), ("entry_point", C.c_uint32), ("graph_entry_point", C.c_uint32),
        ("n_layers", C.c_uint32),
        ("layer_nodes", C.POINTER(c_u32p)), ("layer_offs", C.POINTER(c_u32p)),
        ("layer_nbrs", C.POINTER(c_u32p)), ("layer_sizes", c_u32p),
        ("signs", c_f32p),
    ]


class CpoFlatView(C.Structure):
    _fields_ = [("codes", c_u8p), ("code_stride", C.c_uint64), ("nop_off", C.c_uint32),
                ("ipqo_off", C.c_uint32), ("centroid", c_f32p)]


def code_bytes(D: int, B: int) -> int:
    """sizeof(RaBitQCode<D>) / sizeof(NbitRaBitQCode<D,B>)."""
    words = (D + 63) // 64
    storage = -(-(8 * words * B) // 64) * 64
    return -(-(storage + 8) // 64) * 64


def nb_layout(D: int, B: int) -> dict:
    """Field offsets inside FastScanNeighborBlock / NbitFastScanNeighborBlock (App. B)."""
    o = 4 * D * B
    lay = {"planes": 0, "nop": o, "ip_qo": o + 128, "ip_cp": o + 256, "pop": o + 384}
    if B > 1:
        lay["wpop"] = o + 448
        lay["ids"] = o + 512
    else:
        lay["wpop"] = None
        lay["ids"] = o + 448
    lay["count"] = lay["ids"] + 128
    lay["size"] = -(-(lay["count"] + 4) // 64) * 64
    return lay


class SaveFile:
    """Independent numpy parser of save-file v2 (api/hnsw_index.hpp:217-303)."""

    def __init__(self, path):
        self.path = str(path)
        mm = np.memmap(self.path, dtype=np.uint8, mode="r")
        self.mm = mm
        hdr = bytes(mm[:68])
        magic = int.from_bytes(hdr[0:8], "little")
        version = int.from_bytes(hdr[8:12], "little")
        if magic != MAGIC:
            raise RuntimeError("Invalid magic bytes (not a CP-HNSW index file).")
        if version != 2:
            raise RuntimeError(f"Unsupported index file version: {version}")
        u32 = lambda o: int.from_bytes(hdr[o:o + 4], "little")  # noqa: E731
        self.D, self.R, self.B, self.dim = u32(12), u32(16), u32(20), u32(24)
        self.n = int.from_bytes(hdr[28:36], "little")
        self.max_level = int.from_bytes(hdr[36:40], "little", signed=True)
        self.entry_point = u32(40)
        self.rotation_seed = int.from_bytes(hdr[60:68], "little")
        off = 68
        self.calib_off = off
        cal = bytes(mm[off:off + 248])
        self.calib_bytes = cal
        f32 = lambda o: float(np.frombuffer(cal, np.float32, 1, o)[0])  # noqa: E731
        self.affine_a, self.affine_b, self.ip_qo_floor = f32(0), f32(4), f32(8)
        self.gamma_min, self.gamma_max, self.gamma_beta = f32(80), f32(84), f32(88)
        self.gamma_warmup = int.from_bytes(cal[96:104], "little")
        self.slack_levels = np.frombuffer(cal, np.float32, 32, 108).copy()
        self.num_slack_levels = int.from_bytes(cal[236:240], "little", signed=True)
        self.search_gamma = f32(240)
        off += 248 + 72
        n, D, B, dim = self.n, self.D, self.B, self.dim
        self.centroid = np.frombuffer(mm, np.float32, dim, off).copy(); off += 4 * dim
        self.node_levels = np.frombuffer(mm, np.int32, n, off); off += 4 * n
        self.norm_sq = np.frombuffer(mm, np.float32, n, off); off += 4 * n
        self.raw = np.frombuffer(mm, np.float32, n * D, off).reshape(n, D); off += 4 * n * D
        self.nb_off = code_bytes(D, B)
        self.lay = nb_layout(D, B)
        self.rec_size = self.nb_off + self.lay["size"]
        self.search_data = np.frombuffer(mm, np.uint8, n * self.rec_size, off).reshape(n, self.rec_size)
        off += n * self.rec_size
        n_layers = int(np.frombuffer(mm, np.uint32, 1, off)[0]); off += 4
        self.layers = []
        for _ in range(n_layers):
            ne = int(np.frombuffer(mm, np.uint32, 1, off)[0]); off += 4
            nodes = np.empty(ne, np.uint32)
            offs = np.zeros(ne + 1, np.uint32)
            nbrs = []
            for e in range(ne):
                node, cnt = np.frombuffer(mm, np.uint32, 2, off); off += 8
                nodes[e] = node
                nbrs.append(np.frombuffer(mm, np.uint32, int(cnt), off)); off += 4 * int(cnt)
                offs[e + 1] = offs[e] + cnt
            self.layers.append((nodes, offs, np.concatenate(nbrs).astype(np.uint32) if nbrs else np.zeros(0, np.uint32)))
        self.end = off

    # -- neighbour-block field views (all vertices at once) --------------------------------
    def field(self, name):
        lay, nb = self.lay, self.nb_off
        sd = self.search_data
        if name == "planes":
            return sd[:, nb:nb + 4 * self.D * self.B]
        if name in ("nop", "ip_qo", "ip_cp"):
            o = nb + lay[name]
            return sd[:, o:o + 128].view(np.float32)
        if name in ("pop", "wpop"):
            o = nb + lay[name]
            return sd[:, o:o + 64].view(np.uint16)
        if name == "ids":
            o = nb + lay["ids"]
            return sd[:, o:o + 128].view(np.uint32)
        if name == "count":
            o = nb + lay["count"]
            return sd[:, o:o + 4].view(np.uint32)[:, 0]
        raise KeyError(name)


class Oracle:
    """ctypes handle on the C restatement (+ the reference shim when present)."""

    def __init__(self):
        self.lib = C.CDLL(str(build_port()))
        L = self.lib
        L.cpo_dot.restype = C.c_float
        L.cpo_l2.restype = C.c_float
        L.cpo_dot.argtypes = [C.c_uint32, c_f32p, c_f32p]
        L.cpo_l2.argtypes = [C.c_uint32, c_f32p, c_f32p]
        L.cpo_greedy_descent.restype = C.c_uint32
        self.shim = None
        p = REF_DIR / "libcphnsw_refshim.so"
        if p.exists():
            self.shim = C.CDLL(str(p))
            self.shim.refshim_dot.restype = C.c_float
            self.shim.refshim_l2.restype = C.c_float
            self.shim.refshim_dot.argtypes = [C.c_uint32, c_f32p, c_f32p]
            self.shim.refshim_l2.argtypes = [C.c_uint32, c_f32p, c_f32p]

    # ---- query prep ----------------------------------------------------------------------
    def rotation_signs(self, D, seed=42):
        s = np.empty((3, D), np.float32)
        self.lib.cpo_rotation_signs(C.c_uint32(D), C.c_uint64(seed), _ptr(s, c_f32p))
        return s

    def encode_queries(self, q, D=None, signs=None, want_rotated=False):
        q = np.ascontiguousarray(q, np.float32)
        nq, dim = q.shape
        D = D or (1 << (dim - 1).bit_length())
        signs = self.rotation_signs(D) if signs is None else signs
        lut = np.empty((nq, D // 4, 16), np.uint8)
        co = np.empty((nq, 3), np.float32)
        rot = np.empty((nq, D), np.float32)
        for i in range(nq):
            self.lib.cpo_encode_query(C.c_uint32(dim), C.c_uint32(D), _ptr(signs, c_f32p),
                                      _ptr(q[i], c_f32p), _ptr(lut[i], c_u8p), _ptr(co[i], c_f32p),
                                      _ptr(rot[i], c_f32p))
        return (lut, co, rot) if want_rotated else (lut, co)

    def ref_encode_queries(self, q):
        q = np.ascontiguousarray(q, np.float32)
        nq, dim = q.shape
        D = 1 << (dim - 1).bit_length()
        lut = np.empty((nq, D // 4, 16), np.uint8)
        co = np.empty((nq, 3), np.float32)
        rc = self.shim.refshim_encode_queries(C.c_uint32(dim), C.c_uint64(nq), _ptr(q, c_f32p),
                                              _ptr(lut, c_u8p), _ptr(co, c_f32p))
        assert rc == 0
        return lut, co

    def ref_rotate(self, q):
        q = np.ascontiguousarray(q, np.float32)
        nq, dim = q.shape
        D = 1 << (dim - 1).bit_length()
        out = np.empty((nq, D), np.float32)
        assert self.shim.refshim_rotate(C.c_uint32(dim), C.c_uint64(nq), _ptr(q, c_f32p), _ptr(out, c_f32p)) == 0
        return out

    # ---- build side: neighbour codes relative to a parent (oracle groundwork) --------
    CAQ_FLAGS = 0x1F9          # CPO_CAQ_FLAGS (cphnsw_oracle.h): which a*b+c of the N-bit encoder the reference's compiler fused

    def neighbor_aux(self, dim, B, parent, nbrs, ref=False, fused=None):
        """codes u8 [n, B, D/8] (planes MSB first; B = 1: sign bits) and aux f32 [n, 3] = nop, ip_qo, ip_cp of the
        neighbours `nbrs` [n, dim] of `parent` [dim].  ref=True: the unmodified reference through the shim
        (compute_neighbor_aux / compute_neighbor_aux_nbit); else the C restatement.  `fused` overrides the pinned
        contraction pattern (B = 1: 0/1 for `nop_sq += d * d`; B > 1: the flag word of cpo_neighbor_aux_nbit)."""
        D = 1 << (dim - 1).bit_length()
        par = np.zeros(D, np.float32); par[:dim] = parent
        nb = np.zeros((len(nbrs), D), np.float32); nb[:, :dim] = nbrs
        codes = np.zeros((len(nbrs), B, D // 8), np.uint8)
        aux = np.zeros((len(nbrs), 3), np.float32)
        if ref:
            rc = self.shim.refshim_neighbor_aux(C.c_uint32(D), C.c_uint32(B), C.c_uint32(dim), C.c_uint64(len(nbrs)),
                                                _ptr(par, c_f32p), _ptr(nb, c_f32p), _ptr(codes, c_u8p), _ptr(aux, c_f32p))
            assert rc == 0, "the shim offers D >= 64, B in {1, 2, 4}"
        else:
            signs = self.rotation_signs(D)
            if fused is None:
                fused = 0 if B == 1 else self.CAQ_FLAGS
            for i in range(len(nbrs)):
                if B > 1:
                    self.lib.cpo_neighbor_aux_nbit(C.c_uint32(dim), C.c_uint32(D), C.c_uint32(B), _ptr(signs, c_f32p), _ptr(par, c_f32p),
                                                   _ptr(nb[i], c_f32p), C.c_uint32(fused), _ptr(codes[i], c_u8p), _ptr(aux[i], c_f32p))
                    continue
                self.lib.cpo_neighbor_aux_1bit(C.c_uint32(dim), C.c_uint32(D), _ptr(signs, c_f32p), _ptr(par, c_f32p), _ptr(nb[i], c_f32p),
                                               C.c_int(fused), _ptr(codes[i, 0], c_u8p), _ptr(aux[i], c_f32p))
        return codes, aux

    # ---- fastscan --------------------------------------------------------------------------
    def fastscan(self, D, B, lut, planes, ref=False):
        lut = np.ascontiguousarray(lut, np.uint8)
        planes = np.ascontiguousarray(planes, np.uint8)
        nbit = np.empty(32, np.uint32); msb = np.empty(32, np.uint32); msb2 = np.empty(32, np.uint32)
        fn = self.shim.refshim_fastscan if ref else self.lib.cpo_fastscan
        fn(C.c_uint32(D), C.c_uint32(B), _ptr(lut, c_u8p), _ptr(planes, c_u8p),
           _ptr(nbit, c_u32p), _ptr(msb, c_u32p), _ptr(msb2, c_u32p))
        return nbit, msb, msb2

    def convert(self, D, B, params, nbit, msb, msb2, nop, ip_qo, ip_cp, pop, wpop, count, dqp, ref=False):
        params = np.ascontiguousarray(params, np.float32)
        a = lambda x, t: np.ascontiguousarray(x, t)  # noqa: E731
        nbit, msb, msb2 = a(nbit, np.uint32), a(msb, np.uint32), a(msb2, np.uint32)
        nop, ip_qo, ip_cp = a(nop, np.float32), a(ip_qo, np.float32), a(ip_cp, np.float32)
        pop = a(pop, np.uint16)
        wpop = a(wpop if wpop is not None else np.zeros(32), np.uint16)
        est = np.full(32, np.nan, np.float32); lower = np.full(32, np.nan, np.float32)
        msb_lower = np.full(32, np.nan, np.float32)
        if ref:
            self.shim.refshim_convert(
                C.c_uint32(D), C.c_uint32(B), _ptr(params, c_f32p), _ptr(nbit, c_u32p), _ptr(msb, c_u32p),
                _ptr(msb2, c_u32p), _ptr(nop, c_f32p), _ptr(ip_qo, c_f32p), _ptr(ip_cp, c_f32p),
                _ptr(pop, c_u16p), _ptr(wpop, c_u16p), C.c_uint32(count), C.c_float(dqp),
                _ptr(est, c_f32p), _ptr(lower, c_f32p), _ptr(msb_lower, c_f32p))
        else:
            L = self.lib
            if B == 1:
                L.cpo_convert_1bit(_ptr(params, c_f32p), _ptr(nbit, c_u32p), _ptr(nop, c_f32p), _ptr(ip_qo, c_f32p),
                                   _ptr(ip_cp, c_f32p), _ptr(pop, c_u16p), C.c_uint32(count), C.c_float(dqp),
                                   _ptr(est, c_f32p), _ptr(lower, c_f32p))
                msb_lower[:count] = lower[:count]
            else:
                L.cpo_convert_msb(C.c_uint32(B), _ptr(params, c_f32p), _ptr(msb2, c_u32p), _ptr(nop, c_f32p),
                                  _ptr(ip_qo, c_f32p), _ptr(ip_cp, c_f32p), _ptr(pop, c_u16p), C.c_uint32(count),
                                  C.c_float(dqp), _ptr(msb_lower, c_f32p))
                L.cpo_convert_nbit(C.c_uint32(B), _ptr(params, c_f32p), _ptr(nbit, c_u32p), _ptr(msb, c_u32p),
                                   _ptr(nop, c_f32p), _ptr(ip_qo, c_f32p), _ptr(ip_cp, c_f32p), _ptr(pop, c_u16p),
                                   _ptr(wpop, c_u16p), C.c_uint32(count), C.c_float(dqp),
                                   _ptr(est, c_f32p), _ptr(lower, c_f32p))
        return est, lower, msb_lower

    def dot(self, a, b, ref=False):
        a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
        fn = self.shim.refshim_dot if ref else self.lib.cpo_dot
        return np.float32(fn(C.c_uint32(a.size), _ptr(a, c_f32p), _ptr(b, c_f32p)))

    def l2(self, a, b, ref=False):
        a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
        fn = self.shim.refshim_l2 if ref else self.lib.cpo_l2
        return np.float32(fn(C.c_uint32(a.size), _ptr(a, c_f32p), _ptr(b, c_f32p)))

    # ---- index-level -------------------------------------------------------------------------
    def index_view(self, sf: SaveFile, gamma_override=None):
        """Build a cpo_index over a parsed save file (keeps the arrays alive on the view)."""
        ix = CpoIndex()
        keep = []
        ix.D, ix.B, ix.dim, ix.n = sf.D, sf.B, sf.dim, sf.n
        sd = np.ascontiguousarray(sf.search_data); keep.append(sd)
        raw = np.ascontiguousarray(sf.raw); keep.append(raw)
        ns = np.ascontiguousarray(sf.norm_sq); keep.append(ns)
        ix.search_data = _ptr(sd, c_u8p); ix.rec_size = sf.rec_size; ix.nb_off = sf.nb_off
        ix.raw = _ptr(raw, c_f32p); ix.norm_sq = _ptr(ns, c_f32p)
        ix.affine_a, ix.affine_b, ix.ip_qo_floor = sf.affine_a, sf.affine_b, sf.ip_qo_floor
        for i in range(32):
            ix.slack_levels[i] = float(sf.slack_levels[i])
        ix.num_slack_levels = sf.num_slack_levels
        ix.search_gamma, ix.gamma_max, ix.gamma_beta = sf.search_gamma, sf.gamma_max, sf.gamma_beta
        ix.gamma_warmup = sf.gamma_warmup
        if gamma_override:
            for k_, v in gamma_override.items():
                setattr(ix, k_, v)
        ix.max_level = sf.max_level
        ix.entry_point = sf.entry_point
        ix.graph_entry_point = sf.entry_point  # after load both coincide (hnsw_index.hpp:424-428)
        nl = len(sf.layers)
        ix.n_layers = nl
        nodes_arr = (c_u32p * max(nl, 1))(); offs_arr = (c_u32p * max(nl, 1))(); nbrs_arr = (c_u32p * max(nl, 1))()
        sizes = np.zeros(max(nl, 1), np.uint32)
        for i, (nodes, offs, nbrs) in enumerate(sf.layers):
            nodes = np.ascontiguousarray(nodes); offs = np.ascontiguousarray(offs)
            nbrs = np.ascontiguousarray(nbrs if nbrs.size else np.zeros(1, np.uint32))
            keep += [nodes, offs, nbrs]
            nodes_arr[i] = _ptr(nodes, c_u32p); offs_arr[i] = _ptr(offs, c_u32p); nbrs_arr[i] = _ptr(nbrs, c_u32p)
            sizes[i] = nodes.size
        keep += [nodes_arr, offs_arr, nbrs_arr, sizes]
        ix.layer_nodes = C.cast(nodes_arr, C.POINTER(c_u32p)); ix.layer_offs = C.cast(offs_arr, C.POINTER(c_u32p))
        ix.layer_nbrs = C.cast(nbrs_arr, C.POINTER(c_u32p)); ix.layer_sizes = _ptr(sizes, c_u32p)
        signs = self.rotation_signs(sf.D, sf.rotation_seed); keep.append(signs)
        ix.signs = _ptr(signs, c_f32p)
        ix._keep = keep
        return ix

    # ---- calibration sampling (N4) -----------------------------------------------------------
    CALIB_FLAGS = 3   # contraction of the two scalar expressions as the compiled reference has it (tests/test_oracle_calibration.py)

    @staticmethod
    def _calib_out(ns):
        return {"parent": np.empty(ns, np.uint32), "nn_dist_sq": np.empty(ns, np.float32), "dist_qp_sq": np.empty(ns, np.float32),
                "nop": np.empty((ns, 32), np.float32), "ip_corrected": np.empty((ns, 32), np.float32),
                "ip_qo_denom": np.empty((ns, 32), np.float32), "true_ip": np.empty((ns, 32), np.float32),
                "neighbor": np.empty((ns, 32), np.uint32)}

    def calibration_samples(self, ix, queries_padded, start_ids, flags=None):
        """C restatement of process_query (api/hnsw_index.hpp:786-866) per (query [D], start vertex)."""
        flags = self.CALIB_FLAGS if flags is None else flags
        q = np.ascontiguousarray(queries_padded, np.float32)
        st = np.ascontiguousarray(start_ids, np.uint32)
        assert q.shape == (len(st), ix.D)
        o = self._calib_out(len(st))
        for i in range(len(st)):
            par = C.c_uint32(0); nn = C.c_float(0); dq = C.c_float(0)
            self.lib.cpo_calibration_sample(C.byref(ix), _ptr(q[i], c_f32p), C.c_uint32(int(st[i])), C.c_uint32(flags), C.byref(par), C.byref(nn),
                                            C.byref(dq), _ptr(o["nop"][i], c_f32p), _ptr(o["ip_corrected"][i], c_f32p),
                                            _ptr(o["ip_qo_denom"][i], c_f32p), _ptr(o["true_ip"][i], c_f32p), _ptr(o["neighbor"][i], c_u32p))
            o["parent"][i] = par.value; o["nn_dist_sq"][i] = nn.value; o["dist_qp_sq"][i] = dq.value
        return o

    def ref_calibration_samples(self, sf, queries_padded, start_ids):
        """The same loop composed from the unmodified reference's primitives and block types (oracle/refshim.cpp)."""
        assert self.shim is not None
        q = np.ascontiguousarray(queries_padded, np.float32)
        st = np.ascontiguousarray(start_ids, np.uint32)
        sd = np.ascontiguousarray(sf.search_data); raw = np.ascontiguousarray(sf.raw, np.float32)
        o = self._calib_out(len(st))
        rc = self.shim.refshim_calib_samples(C.c_uint32(sf.D), C.c_uint32(sf.B), C.c_uint32(sf.D), C.c_uint64(len(st)), _ptr(q, c_f32p), _ptr(st, c_u32p),
                                             _ptr(raw, c_f32p), _ptr(sd, c_u8p), C.c_uint64(sf.rec_size), C.c_uint32(sf.nb_off), _ptr(o["parent"], c_u32p),
                                             _ptr(o["nn_dist_sq"], c_f32p), _ptr(o["dist_qp_sq"], c_f32p), _ptr(o["nop"], c_f32p),
                                             _ptr(o["ip_corrected"], c_f32p), _ptr(o["ip_qo_denom"], c_f32p), _ptr(o["true_ip"], c_f32p),
                                             _ptr(o["neighbor"], c_u32p))
        assert rc == 0, "the shim is built for padded dims 32, 128 and 1024"
        return o

    def search_batch(self, ix, queries, k, threads=0):
        q = np.ascontiguousarray(queries, np.float32)
        nq = q.shape[0]
        ids = np.empty((nq, k), np.int64); dists = np.empty((nq, k), np.float32)
        st = CpoStats()
        self.lib.cpo_search_batch(C.byref(ix), _ptr(q, c_f32p), C.c_uint64(nq), C.c_uint64(k),
                                  _ptr(ids, c_i64p), _ptr(dists, c_f32p), C.byref(st), C.c_int(threads))
        return ids, dists, st.as_dict()

    def greedy_descent(self, ix, query):
        qp = np.zeros(ix.D, np.float32); qp[:ix.dim] = query
        nd = C.c_uint64(0)
        return int(self.lib.cpo_greedy_descent(C.byref(ix), _ptr(qp, c_f32p), C.byref(nd)))

    def exhaustive(self, ix, sf: SaveFile, query, k, kprime, id_begin=0, id_end=None):
        assert sf.B == 1, "the exhaustive oracle is defined for the per-vertex 1-bit codes"
        id_end = sf.n if id_end is None else id_end
        fv = CpoFlatView()
        sd = ix._keep[0]
        cen = np.ascontiguousarray(sf.centroid, np.float32)
        fv.codes = _ptr(sd, c_u8p); fv.code_stride = sf.rec_size
        storage = -(-(8 * ((sf.D + 63) // 64)) // 64) * 64
        fv.nop_off, fv.ipqo_off = storage, storage + 4
        fv.centroid = _ptr(cen, c_f32p)
        q = np.ascontiguousarray(query, np.float32)
        ids = np.empty(k, np.uint32); dists = np.empty(k, np.float32)
        m = id_end - id_begin
        sums = np.empty(m, np.uint32); est = np.empty(m, np.float32)
        nout = self.lib.cpo_exhaustive_search(C.byref(ix), C.byref(fv), _ptr(q, c_f32p), C.c_uint64(k),
                                              C.c_uint64(kprime), C.c_uint64(id_begin), C.c_uint64(id_end),
                                              _ptr(ids, c_u32p), _ptr(dists, c_f32p), _ptr(sums, c_u32p),
                                              _ptr(est, c_f32p))
        return ids[:nout], dists[:nout], sums, est


def synthetic(n, dim, seed=1234, clusters=0, sigma_c=4.0):
    """Synthetic base/query generator shared by tests and bench."""
    rng = np.random.default_rng(seed)
    if clusters:
        centers = rng.standard_normal((clusters, dim)).astype(np.float32) * np.float32(sigma_c)
        x = centers[rng.integers(0, clusters, n)] + rng.standard_normal((n, dim)).astype(np.float32)
        return x.astype(np.float32)
    return rng.standard_normal((n, dim)).astype(np.float32)


def build_reference_index(base, bits, path, threads=None):
    """Build + finalize + save with the unmodified reference; returns the reference CPIndex."""
    if threads:
        os.environ["OMP_NUM_THREADS"] = str(threads)
    cph = ref_module()
    idx = cph.CPIndex(dim=base.shape[1], bits=bits)
    idx.build(base)
    idx.finalize()
    idx.save(str(path))
    return idx
