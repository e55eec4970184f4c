"""`CPIndex` -- the reference's Python index class (src/bindings.cpp:115-240) with the query path
on a B200.

Same surface: ``CPIndex(dim, bits=1)``, ``build``, ``finalize``, ``search``, ``search_batch``,
``save``, ``load``, properties ``size``, ``dim``, ``is_finalized``; same argument meaning, return
shapes, padding (-1 / FLT_MAX) and exception types.  Index construction and calibration are not
part of this library: ``build`` / ``finalize`` / ``save`` call the reference's own module
(``cphnsw``: whichever one is importable, or the one given to :func:`set_host_module`), and the
finalized index is handed to the device through the reference's save-file format.  ``load`` and
all searches need only this library.  Searches never run on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import tempfile

import numpy as np

from . import _capi

_host_module = None
_FLT_MAX = np.finfo(np.float32).max
_SUPPORTED_D = (16, 32, 64, 128, 256, 512, 1024, 2048)


def set_host_module(mod) -> None:
    """Use `mod` (an imported reference ``cphnsw`` package) for build / finalize / save."""
    global _host_module
    _host_module = mod


def _host():
    global _host_module
    if _host_module is None:
        try:
            import cphnsw  # the reference package, wherever the user installed it
        except ImportError as e:  # pragma: no cover - depends on the environment
            raise RuntimeError(
                "build/finalize/save are performed by the reference's cphnsw module, which is not "
                "importable here; install it or call cphnsw_b200.set_host_module()") from e
        _host_module = cphnsw
    return _host_module


def _next_pow2(v: int) -> int:
    p = 1
    while p < v:
        p <<= 1
    return p


def _is_torch(x) -> bool:
    return type(x).__module__.split(".")[0] == "torch"


def recover_id_map(save_path: str, vectors, dim: int) -> np.ndarray:
    """internal id -> row of `vectors`, for a save file (api/hnsw_index.hpp:217-303) built from `vectors`.

    finalize() renumbers the vertices in BFS order (graph/rabitq_graph.hpp:208-278) and keeps no map back, but the
    file stores every raw vector in internal order, so rows are matched by their bytes."""
    v = np.ascontiguousarray(np.asarray(vectors), dtype=np.float32)
    with open(save_path, "rb") as f:
        hdr = f.read(68)
    D = int.from_bytes(hdr[12:16], "little")
    fdim = int.from_bytes(hdr[24:28], "little")
    n = int.from_bytes(hdr[28:36], "little")
    if v.ndim != 2 or v.shape[1] != dim or fdim != dim or v.shape[0] != n:
        raise ValueError("vectors must be the (n, dim) float32 array the index was built from")
    off = 68 + 248 + 72 + 4 * dim + 8 * n
    raw = np.memmap(save_path, np.float32, "r", off, (n, D))[:, :dim]
    key = lambda a: np.ascontiguousarray(a).view(np.dtype((np.void, 4 * dim))).ravel()  # noqa: E731
    ko, kr = key(v), key(np.ascontiguousarray(raw))
    order = np.argsort(ko, kind="stable")
    pos = np.minimum(np.searchsorted(ko[order], kr), n - 1)
    m = order[pos].astype(np.uint32)
    if not np.array_equal(ko[m], kr):
        raise ValueError("the index does not hold these vectors")
    return m


class CPIndex:
    """Drop-in for ``cphnsw.CPIndex`` whose ``search`` / ``search_batch`` run on the GPU."""

    def __init__(self, dim: int, bits: int = 1, device: int | None = None):
        dim = int(dim)
        bits = int(bits)
        # the factory's checks and messages (src/bindings.cpp:77-113)
        if dim <= 0 or _next_pow2(dim) not in _SUPPORTED_D:
            raise ValueError(
                f"Unsupported dimension {dim} (padded to {_next_pow2(max(dim, 1))}). Supported padded dims: "
                "16, 32, 64, 128, 256, 512, 1024, 2048.")
        if bits not in (1, 2, 4):
            raise ValueError(f"Unsupported bits={bits}. Supported: 1, 2, 4.")
        self._dim, self._bits = dim, bits
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0")) if "CPHNSW_B200_DEVICE" not in os.environ \
                else int(os.environ["CPHNSW_B200_DEVICE"])
        self._device = device
        self._lib = _capi.lib()
        h = C.c_void_p()
        rc = self._lib.cphnsw_b200_create(device, C.byref(h))
        if rc != _capi.OK:
            _capi.check(None, rc)
        self._h = h
        self._host_index = None   # reference CPIndex (only when built here)
        self._source_path = None  # save file the device copy came from
        self._finalized = False
        self._info = None
        self._tmp = None

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.cphnsw_b200_destroy(self._h)
                self._h = None
            if getattr(self, "_tmp", None):
                self._tmp.cleanup()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass

    # ---- construction: the reference does it ---------------------------------------------------
    def build(self, vectors) -> None:
        v = np.asarray(vectors)
        if v.ndim != 2 or v.shape[1] != self._dim:
            raise ValueError("vectors must be a (n, dim) float32 array")
        v = np.ascontiguousarray(v, dtype=np.float32)
        self._host_index = _host().CPIndex(dim=self._dim, bits=self._bits)
        self._finalized = False
        self._built_from = v
        self._host_index.build(v)

    def finalize(self) -> None:
        if self._host_index is None:
            raise RuntimeError("Finalize called without a pending build.")
        self._host_index.finalize()
        # hand-off through the reference's own serialisation (api/hnsw_index.hpp:217-303)
        self._tmp = tempfile.TemporaryDirectory(prefix="cphnsw_b200_")
        path = os.path.join(self._tmp.name, "index.bin")
        self._host_index.save(path)
        self._load_device(path)
        if getattr(self, "_built_from", None) is not None:
            self.set_original_ids(self._built_from)
            self._built_from = None

    def save(self, path) -> None:
        if not self._finalized:
            raise RuntimeError("Index must be finalized before saving.")
        if self._host_index is not None:
            self._host_index.save(str(path))
        elif os.path.abspath(str(path)) != os.path.abspath(self._source_path):
            shutil.copyfile(self._source_path, str(path))

    def load(self, path) -> None:
        self._host_index = None
        self._id_map = self._id_map_dev = None
        self._load_device(str(path))

    def _load_device(self, path: str) -> None:
        # Index::load validates before it commits anything (api/hnsw_index.hpp:305-443): read the header here, so a
        # file made for another dim / bits is rejected while the device still holds the old index
        try:
            with open(path, "rb") as f:
                hdr = f.read(68)
        except OSError as e:
            raise RuntimeError(f"Cannot open file for reading: {path}") from e
        if len(hdr) == 68 and int.from_bytes(hdr[0:8], "little") == 0x57534E48504300 and int.from_bytes(hdr[8:12], "little") == 2:
            fbits, fdim = int.from_bytes(hdr[20:24], "little"), int.from_bytes(hdr[24:28], "little")
            if fdim != self._dim or fbits != self._bits:
                raise RuntimeError(
                    f"Parameter mismatch: file has dim={fdim}, bits={fbits}; index was created with "
                    f"dim={self._dim}, bits={self._bits}.")
        # (anything else -- magic, version, truncation -- is the native loader's to report, with the reference's messages)
        rc = self._lib.cphnsw_b200_load(self._h, path.encode())
        info = _capi.Info()
        still = self._lib.cphnsw_b200_get_info(self._h, C.byref(info)) == _capi.OK
        if rc != _capi.OK:
            # the native loader validates the file before it lets go of the old index; if it failed later (during
            # the upload) nothing is left on the device and the object says so
            if not still:
                self._finalized, self._info = False, None
            _capi.check(self._h, rc)
        self._info = info
        self._source_path = path
        self._finalized = True

    # ---- the hot path --------------------------------------------------------------------------
    def search(self, query, k: int = 10):
        q = np.asarray(query.detach().cpu() if _is_torch(query) else query)
        if q.ndim != 1 or q.shape[0] != self._dim:
            raise ValueError("query must be 1D and match index dimension")
        kk = max(int(k), 1)   # Index::search searches with max(k,1) and returns what it found
        ids, dists = self.search_batch(q[None, :], kk)
        m = int((ids[0] >= 0).sum())
        return ids[0, :m].copy(), dists[0, :m].copy()

    def search_batch(self, queries, k: int = 10):
        """(ids int64 [nq,k], distances float32 [nq,k]); rows padded with -1 / FLT_MAX.

        numpy (or CPU torch) input -> numpy output, host<->device copies inside the call.
        CUDA torch tensor input -> CUDA torch tensors, no host round trip.
        """
        k = int(k)
        if k < 0:
            raise ValueError("k must be non-negative")
        if _is_torch(queries) and queries.is_cuda:
            return self._search_batch_cuda(queries, k)
        q = np.asarray(queries.detach() if _is_torch(queries) else queries)
        if q.ndim != 2 or q.shape[1] != self._dim:
            raise ValueError("queries must be a (n, dim) array")
        q = np.ascontiguousarray(q, dtype=np.float32)
        nq = q.shape[0]
        ids = np.empty((nq, k), np.int64)
        dists = np.empty((nq, k), np.float32)
        self._require_finalized()
        _capi.check(self._h, self._lib.cphnsw_b200_search_batch(
            self._h, q.ctypes.data, nq, k, ids.ctypes.data, dists.ctypes.data))
        return ids, dists

    # ---- the same call in two halves: keep two batches in flight ---------------------------------
    def search_batch_submit(self, queries, k: int = 10, out=None):
        """Enqueue `search_batch(queries, k)` and return a ticket; `search_batch_wait(ticket)` returns the
        (ids, distances) numpy arrays.  Two batches may be in flight per index: the drain of one batch's kernel
        then overlaps the start of the next (that tail is ~15 % of a 10k-query batch).  `queries` may be a numpy
        array or a CPU torch tensor -- page-locked memory (``tensor.pin_memory()``) makes the copies asynchronous;
        `out` = (ids, dists) numpy arrays / CPU tensors to fill instead of fresh ones."""
        k = int(k)
        if k < 0:
            raise ValueError("k must be non-negative")
        q = queries.detach().numpy() if _is_torch(queries) else np.asarray(queries)
        if q.ndim != 2 or q.shape[1] != self._dim:
            raise ValueError("queries must be a (n, dim) array")
        q = np.ascontiguousarray(q, dtype=np.float32)
        nq = q.shape[0]
        if out is None:
            ids, dists = np.empty((nq, k), np.int64), np.empty((nq, k), np.float32)
        else:
            ids, dists = (o.numpy() if _is_torch(o) else o for o in out)
            if ids.shape != (nq, k) or dists.shape != (nq, k) or ids.dtype != np.int64 or dists.dtype != np.float32 \
                    or not ids.flags.c_contiguous or not dists.flags.c_contiguous:
                raise ValueError("out must be C-contiguous (int64 [nq,k], float32 [nq,k]) arrays")
        self._require_finalized()
        t = C.c_uint64(0)
        _capi.check(self._h, self._lib.cphnsw_b200_search_batch_submit(
            self._h, q.ctypes.data, nq, k, ids.ctypes.data, dists.ctypes.data, C.byref(t)))
        self._pending = getattr(self, "_pending", {})
        self._pending[t.value] = (q, ids, dists)   # keeps the buffers alive until the wait
        return t.value

    def search_batch_wait(self, ticket: int):
        q_ids_dists = getattr(self, "_pending", {}).pop(int(ticket), None)
        if q_ids_dists is None:
            raise ValueError("unknown ticket")
        _capi.check(self._h, self._lib.cphnsw_b200_search_batch_wait(self._h, int(ticket)))
        return q_ids_dists[1], q_ids_dists[2]

    def search_batches(self, batches, k: int = 10):
        """Generator over (ids, distances) of each array in `batches`, with two batches in flight."""
        prev = None
        for b in batches:
            t = self.search_batch_submit(b, k)
            if prev is not None:
                yield self.search_batch_wait(prev)
            prev = t
        if prev is not None:
            yield self.search_batch_wait(prev)

    def synchronize(self) -> None:
        """Wait for every call in flight on this index (device-tensor searches are asynchronous)."""
        _capi.check(self._h, self._lib.cphnsw_b200_synchronize(self._h))

    def _search_batch_cuda(self, queries, k: int):
        import torch

        if queries.dim() != 2 or queries.shape[1] != self._dim:
            raise ValueError("queries must be a (n, dim) array")
        if queries.device.index != self._device:
            raise ValueError(f"queries live on cuda:{queries.device.index}, the index on cuda:{self._device}")
        q = queries.detach().to(torch.float32).contiguous()
        nq = q.shape[0]
        ids = torch.empty((nq, k), dtype=torch.int64, device=q.device)
        dists = torch.empty((nq, k), dtype=torch.float32, device=q.device)
        self._require_finalized()
        stream = torch.cuda.current_stream(q.device).cuda_stream
        _capi.check(self._h, self._lib.cphnsw_b200_search_batch_device(
            self._h, q.data_ptr(), nq, k, ids.data_ptr(), dists.data_ptr(), stream))
        return ids, dists

    # ---- opt-in clean-up of the reference's result conventions (outside the parity path) ---------
    def search_batch_unique(self, queries, k: int = 10, k_search: int | None = None, original_ids: bool = False):
        """`search_batch(queries, k_search)` -- bit-identical to the reference's -- followed on the device by
        de-duplication (the reference lists a vertex once per time it was scored, rabitq_search.hpp:133,236,250)
        and, with ``original_ids=True``, translation of the internal BFS-reordered ids
        (rabitq_graph.hpp:208-278) to row numbers of the array given to `build` / `set_original_ids`.

        Returns the k closest distinct neighbours found, padded with -1 / FLT_MAX.  k_search defaults to 3k
        (the reference returns ~5.3 distinct ids out of 10 on iid data)."""
        import torch

        k = int(k)
        ks = int(k_search) if k_search is not None else 3 * max(k, 1)
        if ks < k:
            raise ValueError("k_search must be at least k")
        dev = torch.device("cuda", self._device)
        cuda_in = _is_torch(queries) and queries.is_cuda
        q = queries if cuda_in else torch.from_numpy(np.ascontiguousarray(np.asarray(
            queries.detach() if _is_torch(queries) else queries), dtype=np.float32)).to(dev)
        ids, dists = self._search_batch_cuda(q, ks)
        out_i = torch.empty((q.shape[0], k), dtype=torch.int64, device=dev)
        out_d = torch.empty((q.shape[0], k), dtype=torch.float32, device=dev)
        idmap = self._device_id_map() if original_ids else None
        _capi.check(self._h, self._lib.cphnsw_b200_unique_topk(
            self._h, ids.data_ptr(), dists.data_ptr(), q.shape[0], ks, k,
            idmap.data_ptr() if idmap is not None else None, idmap.numel() if idmap is not None else 0,
            out_i.data_ptr(), out_d.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        if cuda_in:
            return out_i, out_d
        return out_i.cpu().numpy(), out_d.cpu().numpy()

    def set_original_ids(self, vectors) -> None:
        """Recover the internal-id -> original-row map by matching the index's stored vectors (save-file
        section `raw`, internal order: api/hnsw_index.hpp:217-303) against `vectors` (the array the index was
        built from).  Rows of `vectors` that are bit-identical are indistinguishable: any of them is returned."""
        self._require_finalized()
        self._id_map = recover_id_map(self._source_path, vectors, self._dim)
        self._id_map_dev = None

    def _device_id_map(self):
        import torch

        if getattr(self, "_id_map", None) is None:
            raise RuntimeError("original ids are unknown: call set_original_ids(vectors) first (build() does it)")
        if getattr(self, "_id_map_dev", None) is None:
            self._id_map_dev = torch.from_numpy(self._id_map.view(np.int32)).to(torch.device("cuda", self._device))
        return self._id_map_dev

    def _require_finalized(self):
        if not self._finalized:
            raise RuntimeError("Index must be finalized (or loaded) before searching on the device.")

    # ---- introspection -------------------------------------------------------------------------
    @property
    def size(self) -> int:
        if self._info is not None:
            return int(self._info.n)
        return int(self._host_index.size) if self._host_index is not None else 0

    @property
    def dim(self) -> int:
        return self._dim

    @property
    def is_finalized(self) -> bool:
        return self._finalized

    @property
    def handle(self):
        """Opaque native handle (for the kernel-level hooks in cphnsw_b200.hooks)."""
        return self._h

    def info(self) -> dict:
        self._require_finalized()
        i = self._info
        d = {n: getattr(i, n) for n, _ in i._fields_ if n != "slack_levels"}
        d["slack_levels"] = [float(x) for x in i.slack_levels][: max(i.num_slack_levels, 0)]
        return d

    def last_stats(self) -> dict:
        st = _capi.Stats()
        _capi.check(self._h, self._lib.cphnsw_b200_last_stats(self._h, C.byref(st)))
        return st.as_dict()

    def last_timings(self) -> dict:
        """Device milliseconds of the last search's kernels (CUDA events on the launching stream)."""
        a, b = C.c_float(0), C.c_float(0)
        _capi.check(self._h, self._lib.cphnsw_b200_last_timings(self._h, C.byref(a), C.byref(b)))
        return {"prep_ms": a.value, "search_ms": b.value}

    def set_option(self, name: str, value: int) -> None:
        _capi.check(self._h, self._lib.cphnsw_b200_set_option(self._h, name.encode(), int(value)))
