#!/usr/bin/env python
"""bench.py -- QPS of the CP-HNSW query path (RaBitQ FastScan estimate + beam search + exact-L2 rerank).

    python bench.py --gpus N --steps K --warmup W                 # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference's CPU path (oracle/_ref)

One step = one search_batch over the workload's query batch (BASELINE.json configs[1]: 1M x 128,
4-bit codes, 10k queries, k = 10).  The index is built ONCE by the reference's own build code
(oracle/_ref/refbuild, see its header for why the stock finalize() cannot do 1M) and both arms load
that same file; it is looked up in bench_cache/ (prebuilt, travels with the tree), then in
/tmp/cphnsw_b200_bench_cache/, and built on the spot otherwise.  N > 1: one process per GPU, index
replicated, each rank searches its own 10k queries (weak scaling, no data-path collective).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
T_START = time.time()
sys.path.insert(0, str(ROOT / "rabitq-ann-search_b200"))
TMP_CACHE = Path("/tmp/cphnsw_b200_bench_cache")
# The reference's NNDescent build depends on its OpenMP thread count (16 threads and 32 threads give graphs that cost
# 2 801 and 4 103 expansions per query at 1M x 128 x 4-bit); pinned, so that every box searches the same graph.
BUILD_THREADS = 16
FLT_MAX = np.finfo(np.float32).max


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------------
def synthetic(n, dim, seed, clusters=0, sigma_c=4.0):
    rng = np.random.default_rng(seed)
    if clusters:
        centers = rng.standard_normal((clusters, dim)).astype(np.float32) * np.float32(sigma_c)
        return (centers[rng.integers(0, clusters, n)] + rng.standard_normal((n, dim)).astype(np.float32)).astype(np.float32)
    return rng.standard_normal((n, dim)).astype(np.float32)


def make_queries(args, rank):
    if args.clusters:
        centers = np.random.default_rng(args.seed).standard_normal((args.clusters, args.dim)).astype(np.float32) * np.float32(4.0)
        rng = np.random.default_rng(99 + rank)
        return (centers[rng.integers(0, args.clusters, args.nq)] + rng.standard_normal((args.nq, args.dim)).astype(np.float32)).astype(np.float32)
    return np.random.default_rng(99 + rank).standard_normal((args.nq, args.dim)).astype(np.float32)


def index_name(args):
    dist = f"cl{args.clusters}" if args.clusters else "iid"
    return f"c2_{args.n}x{args.dim}_b{args.bits}_{dist}_seed{args.seed}.bin"


def obtain_index(args, rank, world):
    """Path of the finalized index file, building it with the reference's own code if nobody has yet."""
    name = index_name(args)
    for d, src in ((ROOT / "bench_cache", "prebuilt (bench_cache/, built by oracle/_ref/refbuild)"), (TMP_CACHE, "cached in /tmp")):
        if (d / name).exists():
            return d / name, src
    TMP_CACHE.mkdir(parents=True, exist_ok=True)
    path = TMP_CACHE / name
    if rank == 0:
        refbuild = ROOT / "oracle" / "_ref" / "refbuild"
        if not refbuild.exists():
            raise SystemExit("no index file and no oracle/_ref/refbuild to build one (run __graft_entry__.build() where /root/reference exists)")
        threads = min(BUILD_THREADS, os.cpu_count() or 1)
        log(f"[bench] building {name} with the reference's build code (OMP_NUM_THREADS={threads} of {os.cpu_count()} cores) ...")
        vec = TMP_CACHE / (name + ".f32")
        synthetic(args.n, args.dim, args.seed, args.clusters).tofile(vec)
        t = time.time()
        tmp = str(path) + ".part"
        env = dict(os.environ, OMP_NUM_THREADS=str(threads))   # (torchrun exports OMP_NUM_THREADS=1 to its workers)
        subprocess.run([str(refbuild), str(args.dim), str(args.bits), str(args.n), str(vec), tmp], check=True, stdout=sys.stderr, env=env)
        os.replace(tmp, path)
        vec.unlink()
        log(f"[bench] index built in {time.time() - t:.0f} s")
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
    return path, "built on this box by oracle/_ref/refbuild"


def index_fingerprint(path):
    """Hash of the index file's header, calibration and 64 evenly spaced 1 MiB windows: cheap on a 3 GB file, and both arms
    print it, so a bench line says which graph it searched."""
    import hashlib

    h = hashlib.blake2b(digest_size=8)
    size = os.path.getsize(path)
    with open(path, "rb") as f:
        h.update(f.read(68 + 248 + 72))
        for i in range(64):
            f.seek((size // 64) * i)
            h.update(f.read(1 << 20))
    h.update(str(size).encode())
    return h.hexdigest()


def run_config(args, path):
    """The `config` object: identical for both arms (the driver compares them)."""
    dist = "clustered(%d)" % args.clusters if args.clusters else "iid N(0,1)"
    return {"workload": f"{args.n}x{args.dim} synthetic {dist}, {args.bits}-bit RaBitQ CP-HNSW graph search, {args.nq} queries/GPU, k={args.k}",
            "index_file": index_name(args), "index_fingerprint": index_fingerprint(path),
            "index_source": f"built by the reference's own build code (oracle/_ref/refbuild over the unmodified headers, OMP_NUM_THREADS={BUILD_THREADS})",
            "calibration": "stock finalize() up to n = 230 400; beyond that the reference's EVT fit cannot converge (adaptive_defaults.hpp:45 vs "
                           "evt_crc.hpp:216-233) and refbuild re-runs the reference's calibrate_estimator with evt_min_tail relaxed to what its "
                           "sample supports -- both arms load the same file"}


def parse_save_header(path):
    with open(path, "rb") as f:
        h = f.read(68)
    u32 = lambda o: int.from_bytes(h[o:o + 4], "little")  # noqa: E731
    return {"D": u32(12), "bits": u32(20), "dim": u32(24), "n": int.from_bytes(h[28:36], "little")}


def raw_vectors(path):
    """Raw vectors in internal-id order straight from the save file -- for ground truth."""
    h = parse_save_header(path)
    off = 68 + 248 + 72 + 4 * h["dim"] + 8 * h["n"]
    return np.memmap(path, np.float32, "r", off, (h["n"], h["D"]))


def same_results(ids_a, d_a, ids_b, d_b):
    """(ids equal, distance bits equal) as multisets per row: rows ordered by (distance bits, id) first -- the order of
    equal distances inside a row is unspecified in the reference (sort_heap)."""
    def canon(i, d):
        i = np.asarray(i, np.int64)
        b = np.ascontiguousarray(np.asarray(d, np.float32)).view(np.uint32).astype(np.int64)
        o = np.lexsort((i, b), axis=1)
        return np.take_along_axis(i, o, 1), np.take_along_axis(b, o, 1)
    ia, ba = canon(ids_a, d_a)
    ib, bb = canon(ids_b, d_b)
    return bool(np.array_equal(ia, ib)), bool(np.array_equal(ba, bb))


def recall_at_k(ids, gt):
    """Reference definition (cphnsw/eval.py:23-28): |set(result) & set(truth)| / k, averaged."""
    k = gt.shape[1]
    hit = 0
    for r, g in zip(ids, gt):
        hit += len(set(r[:k].tolist()) & set(g.tolist()))
    return hit / (len(ids) * k)


def ground_truth(path, q, k, device):
    import torch

    base = raw_vectors(path)
    dim = q.shape[1]
    qt = torch.from_numpy(q).to(device)
    best_d = torch.full((q.shape[0], k), float("inf"), device=device)
    best_i = torch.zeros((q.shape[0], k), dtype=torch.int64, device=device)
    step = 262144
    for s in range(0, base.shape[0], step):
        b = torch.from_numpy(np.array(base[s:s + step, :dim])).to(device)
        d = (qt * qt).sum(1, keepdim=True) - 2.0 * (qt @ b.T) + (b * b).sum(1)[None, :]
        dd, ii = torch.topk(d, k, dim=1, largest=False)
        cat_d = torch.cat([best_d, dd], 1)
        cat_i = torch.cat([best_i, ii + s], 1)
        best_d, sel = torch.topk(cat_d, k, dim=1, largest=False)
        best_i = torch.gather(cat_i, 1, sel)
    return best_i.cpu().numpy()


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, sampled in-process through NVML (spawning
    nvidia-smi in a loop stalls the GPU for tens of ms per query and would distort the measurement)."""

    def __init__(self, gpu, period=0.02):
        self.gpu, self.period, self.rows, self.stop = gpu, period, [], False
        self.t = None

    def __enter__(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[self.gpu].isdigit() else self.gpu
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        except Exception as e:  # noqa: BLE001
            log(f"[bench] NVML clock sampling unavailable: {e}")
        return self

    def _run(self):
        nv = self.nv
        while not self.stop:
            try:
                self.rows.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h),
                                  nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0, time.time()))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def __exit__(self, *a):
        self.stop = True
        if self.t:
            self.t.join(timeout=2)

    def window(self, t0, t1):
        """Keep only the samples taken inside [t0, t1] (the timed region)."""
        inside = [r for r in self.rows if t0 <= r[3] <= t1]
        if inside:
            self.rows = inside

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        seen = sorted(n for n, bit in names.items() if any(r[1] & bit for r in self.rows))
        return {"sm_mhz": float(np.median([r[0] for r in self.rows])), "sm_max_mhz": float(self.max), "reasons": seen,
                "samples": len(self.rows), "power_w_max": max(r[2] for r in self.rows)}


# ---------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the unmodified reference's search_batch on the host cores
# ---------------------------------------------------------------------------------------------------
def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def reference_one_thread(args, path, k, budget_s=5.0):
    """QPS of the reference's search_batch with ONE OpenMP thread, in a subprocess because libgomp reads
    OMP_NUM_THREADS once."""
    code = (
        "import sys, time, json, numpy as np; sys.path.insert(0, %r); import cphnsw\n"
        "ix = cphnsw.CPIndex(dim=%d, bits=%d); ix.load(%r)\n"
        "q = np.random.default_rng(99).standard_normal((4096, %d)).astype(np.float32)\n"
        "t = time.perf_counter(); ix.search_batch(q[:8], %d); per = (time.perf_counter() - t) / 8\n"
        "m = int(max(8, min(4096, %f / max(per, 1e-9))))\n"
        "t = time.perf_counter(); ix.search_batch(q[:m], %d); dt = time.perf_counter() - t\n"
        "print(json.dumps({'value': m / dt, 'queries': m}))\n"
    ) % (str(ROOT / "oracle" / "_ref"), args.dim, args.bits, str(path), args.dim, k, budget_s, k)
    try:
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120, env=dict(os.environ, OMP_NUM_THREADS="1"))
        return json.loads(r.stdout.strip().split("\n")[-1])
    except Exception as e:  # noqa: BLE001
        log(f"[bench] 1-thread reference run skipped: {e}")
        return None


def reference_module():
    # all host cores for the reference's OpenMP loop (src/bindings.cpp:196-200), also under torchrun, which
    # exports OMP_NUM_THREADS=1 to its workers; must be set before libgomp initialises
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count())
    ref_dir = ROOT / "oracle" / "_ref"
    if not any((ref_dir / "cphnsw").glob("_core*.so")):
        return None
    sys.path.insert(0, str(ref_dir))
    import cphnsw

    return cphnsw


def time_reference(args, path, q, budget_s, steps, warmup, k=None):
    """QPS of the stock reference CPIndex.search_batch (all host cores) on a bounded sample of q."""
    k = args.k if k is None else k
    cph = reference_module()
    if cph is None:
        return None
    cores = os.cpu_count()
    idx = cph.CPIndex(dim=args.dim, bits=args.bits)
    t = time.time()
    idx.load(str(path))
    log(f"[bench] reference loaded the index in {time.time() - t:.1f} s")
    probe = min(len(q), 4 * cores)
    t = time.perf_counter()
    idx.search_batch(q[:probe], k)
    per_q = (time.perf_counter() - t) / probe
    m = int(min(len(q), max(probe, budget_s / max(per_q, 1e-9) / max(steps + warmup, 1))))
    sample = q[:m]
    for _ in range(warmup):
        idx.search_batch(sample, k)
    times = []
    for _ in range(steps):
        t = time.perf_counter()
        ids, dists = idx.search_batch(sample, k)
        times.append(time.perf_counter() - t)
    total = sum(times)
    return {"value": m * steps / total, "unit": "queries/s", "cores": cores, "kind": "reference",
            "sample": f"first {m} of the {len(q)} queries x {steps} steps, CPIndex.search_batch from oracle/_ref (unmodified reference, OMP all cores)",
            "ms_per_step": 1e3 * total / steps, "ids": ids, "dists": dists, "m": m}


# ---------------------------------------------------------------------------------------------------
# the operating point the metric names: recall@10 >= gate
# ---------------------------------------------------------------------------------------------------
def recall_sweep(args, local, torch, cph, n, bits, ks_list, ix=None, timed=True):
    """Graph search on the n x dim index with `bits`-bit codes: search k_search results exactly as the reference does
    (ids identical to the reference's), de-duplicate on the device, keep k; recall@k against brute force and QPS at each
    k_search of ks_list, stopping at the first that clears the gate.  At the gate both arms are timed (this arm device-
    resident and end to end; the reference on a bounded sample of the queries, de-duplication in numpy untimed) and
    their ids compared.  The reference's k results hold ~5 distinct ids (it lists a vertex once per time it was scored,
    SURVEY F2), which caps recall@10 near 0.5 at k_search = k for both arms: hence k_search > k."""
    import copy

    ga = copy.copy(args)
    ga.n, ga.bits = n, bits
    path, src = obtain_index(ga, 0, 1)
    if ix is None:
        ix = cph.CPIndex(ga.dim, ga.bits, device=local)
        ix.load(str(path))
    q = make_queries(ga, 0)
    q_dev = torch.from_numpy(q).cuda()
    gt = ground_truth(path, q, ga.k, torch.device("cuda", local))
    out = {"workload": f"{ga.n}x{ga.dim} synthetic iid N(0,1), {ga.bits}-bit RaBitQ CP-HNSW graph search, {ga.nq} queries, k={ga.k}",
           "bits": bits, "n": n, "index_file": index_name(ga), "index_origin": src, "gate": args.gate, "curve": [], "reached": False}
    for ks in ks_list:
        ix.search_batch_unique(q_dev, ga.k, k_search=ks)      # warm: buffers for this k
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ui, _ = ix.search_batch_unique(q_dev, ga.k, k_search=ks)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        rec = recall_at_k(ui.cpu().numpy(), gt)
        out["curve"].append({"k_search": ks, "recall_at_10": rec, "qps": ga.nq / dt})
        log(f"[bench] {n}x{ga.dim} {bits}-bit: k_search={ks}: de-duplicated recall@{ga.k} = {rec:.4f} at {ga.nq / dt:.0f} QPS")
        if rec >= args.gate:
            out.update({"reached": True, "k_search": ks, "recall_at_10": rec})
            break
    if not out["reached"] or not timed:
        return out
    ks = out["k_search"]
    steps = max(1, min(args.steps, 5))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(steps):
        ix.search_batch_unique(q_dev, ga.k, k_search=ks)
    ev[1].record()
    torch.cuda.synchronize()
    out["value"] = ga.nq * steps / (ev[0].elapsed_time(ev[1]) / 1e3)
    t0 = time.perf_counter()
    for _ in range(steps):
        hi, _ = ix.search_batch_unique(q, ga.k, k_search=ks)      # numpy in, numpy out
    out["e2e"] = ga.nq * steps / (time.perf_counter() - t0)
    out["unit"] = "queries/s"
    out["how"] = (f"search_batch(k={ks}) as the reference does it, then the first {ga.k} distinct ids (cphnsw_b200_unique_topk on the device); "
                  "one batch at a time")
    if not args.no_cpu_baseline:
        r = time_reference(ga, path, q, budget_s=min(args.cpu_budget, 10.0), steps=1, warmup=0, k=ks)
        if r is not None:
            ri = np.full((r["m"], ga.k), -1, np.int64)
            for row, srcrow in enumerate(r["ids"]):
                u = list(dict.fromkeys(int(x) for x in srcrow if x >= 0))[:ga.k]
                ri[row, :len(u)] = u
            out["cpu_baseline"] = {"value": r["value"], "unit": "queries/s", "cores": r["cores"], "kind": "reference",
                                   "sample": r["sample"] + f" at k={ks}; de-duplication (numpy) not timed",
                                   "recall_at_10": recall_at_k(ri, gt[:r["m"]]),
                                   "ids_identical_to_gpu": bool(np.array_equal(np.sort(ri, 1), np.sort(hi[:r["m"]], 1)))}
            out["speedup_vs_cpu_baseline"] = out["e2e"] / r["value"]
    return out


# ---------------------------------------------------------------------------------------------------
# workload c4 (BASELINE config 4): exhaustive batched scan over 1-bit per-vertex codes, database sharded over the ranks
# ---------------------------------------------------------------------------------------------------
def _c4_calibration() -> bytes:
    """A CalibrationSnapshot (api/hnsw_index.hpp:33-58) as finalize() leaves it in practice: a = 1, b = 0."""
    import struct

    buf = bytearray(248)
    struct.pack_into("<3f", buf, 0, 1.0, 0.0, 0.3)
    struct.pack_into("<3f", buf, 80, 1.0e7, 1.0e7, 0.0)
    struct.pack_into("<f", buf, 240, 1.0e7)
    return bytes(buf)


def _c4_upload(cph, hooks, dim, D, local, sd, raw, norm_sq, centroid):
    ix = cph.CPIndex(dim, 1, device=local)
    hooks.upload_arrays(ix, D=D, bits=1, dim=dim, search_data=sd, raw=raw, norm_sq=norm_sq, calibration=_c4_calibration(),
                        centroid=centroid, max_level=0, entry_point=0, graph_entry_point=0, rotation_seed=42, layers=[])
    return ix


def run_c4(args, rank, world, local):
    """n x 96 (padded to 128) synthetic vectors, 1-bit RaBitQ codes, nq queries, top-k' estimates re-ranked exactly.
    The reference has no exhaustive mode and cannot build 10M vectors in bench time, and the scan reads only the
    per-vertex codes, so each rank encodes its own shard exactly as RaBitQEncoder::encode_impl does
    (encoder/rabitq_encoder.hpp:225-262; the rotation is K1's, i.e. the index's own) and leaves the neighbour
    blocks empty.  Step = scan of the rank's shard for all queries, NCCL all-gather of the per-shard top-k,
    device-side k-way merge.  Strong scaling: the database is fixed, shards shrink with N."""
    import torch
    import torch.distributed as dist

    import cphnsw_b200 as cph
    from cphnsw_b200 import hooks, sharding

    dev = torch.device("cuda", local)
    n, dim, D, k, kp, nq = args.n, 96, 128, args.k, args.kprime, args.nq
    b, e = sharding.db_shard(n, rank, world)
    m = e - b
    chunk = 1 << 20

    def gen_chunk(c):   # chunk c of the database, the same on every rank
        g = torch.Generator(device=dev); g.manual_seed(args.seed + c)
        return torch.randn((min(chunk, n - c * chunk), dim), generator=g, device=dev)

    t0 = time.time()
    base = torch.empty((m, dim), dtype=torch.float32, device=dev)
    for c in range(b // chunk, (e + chunk - 1) // chunk):
        x = gen_chunk(c)
        lo, hi = max(b, c * chunk), min(e, c * chunk + x.shape[0])
        base[lo - b:hi - b] = x[lo - c * chunk:hi - c * chunk]
    csum = base.sum(0, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(csum)
    centroid = (csum / n).to(torch.float32)
    cen_np = centroid.cpu().numpy()
    # record layout of VertexSearchData<128,32,1>: per-vertex code at 0 (signs 16 B, nop @64, ip_qo @68),
    # neighbour block at 128: ids @ +960 (0xFFFFFFFF = empty), count @ +1088
    rec, nb_off = 1280, 128
    tiny = _c4_upload(cph, hooks, dim, D, local, np.zeros((64, rec), np.uint8), np.zeros((64, D), np.float32), np.zeros(64, np.float32), cen_np)
    sd = np.zeros((m, rec), np.uint8)
    raw = np.zeros((m, D), np.float32)
    norm_sq = np.empty(m, np.float32)
    step = 1 << 18
    shifts = torch.arange(8, device=dev, dtype=torch.uint8)
    for s in range(0, m, step):
        x = base[s:s + step]
        rot = hooks.prepare_queries(tiny, x, center=True)["rotated"]
        nop = (x - centroid).norm(dim=1)
        packed = ((rot >= 0).to(torch.uint8).reshape(-1, D // 8, 8) << shifts).sum(2).to(torch.uint8)
        ipqo = (rot.abs().sum(1) / (nop.clamp_min(1e-30) * (D ** 0.5))).to(torch.float32)
        t = s + x.shape[0]
        sd[s:t, :D // 8] = packed.cpu().numpy()
        sd[s:t, 64:68] = nop.cpu().numpy().view(np.uint8).reshape(-1, 4)
        sd[s:t, 68:72] = ipqo.cpu().numpy().view(np.uint8).reshape(-1, 4)
        sd[s:t, nb_off + 960:nb_off + 1088] = 0xFF
        raw[s:t, :dim] = x.cpu().numpy()
        norm_sq[s:t] = (x * x).sum(1).cpu().numpy()
    del tiny
    ix = _c4_upload(cph, hooks, dim, D, local, sd, raw, norm_sq, cen_np)
    del sd, raw
    ix.set_option("exhaustive_tensor_cores", args.scan_form)
    info = ix.info()
    log(f"[bench] rank {rank}: shard [{b}, {e}) encoded and on cuda:{local} in {time.time() - t0:.1f} s, {info['device_bytes'] / 2**30:.1f} GiB")

    gq = torch.Generator(device=dev); gq.manual_seed(99)
    q_dev = torch.randn((nq, dim), generator=gq, device=dev)
    q_pin = q_dev.cpu().pin_memory()

    def step_dev(q, kp_=None, timeline=None):
        kp_ = kp if kp_ is None else kp_
        if world == 1 and not args.c4_pieces:
            return hooks.exhaustive_search(ix, q, k, kp_)           # one scan of the whole database
        # shards scan their ranges in pieces, exchange thresholds (all-reduce(min) of nq floats), all-gather their k' candidates
        # (key + exact distance) and merge: the same result as the single scan (tests/test_exhaustive_gpu.py, test_multigpu_gpu.py)
        return sharding.exhaustive_search_db_sharded_device(ix, q, k, kp_, m, n, b, prefix=args.c4_prefix, growth=args.c4_growth, timeline=timeline)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t_ = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        return float(t_.item())

    def timed(kp_, steps, warmup):
        evs_ = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        for _ in range(warmup):
            step_dev(q_dev, kp_)
        barrier()
        t0_ = time.time()
        evs_[0].record()
        for i in range(steps):
            out = step_dev(q_dev, kp_)
            evs_[i + 1].record()
        barrier()
        return max_over_ranks(evs_[0].elapsed_time(evs_[-1])), [evs_[i].elapsed_time(evs_[i + 1]) for i in range(steps)], out, t0_

    with ClockSampler(local) as clocks:
        dev_ms, step_ms, (out_i, out_d), t_begin = timed(kp, args.steps, args.warmup)
        clocks.window(t_begin, time.time())
    value = nq * args.steps / (dev_ms / 1e3)
    phases = None
    if world > 1 or args.c4_pieces:      # where one step's time goes on this rank (device timeline of one more step)
        tl = []
        step_dev(q_dev, kp, tl)
        torch.cuda.synchronize()
        phases = [{"phase": name, "ms": round(tl[i - 1][1].elapsed_time(ev), 3)} for i, (name, ev) in enumerate(tl) if i]
    # e2e: queries from pinned host memory, results back to the host, every step
    ids_pin = torch.empty((nq, k), dtype=torch.int64).pin_memory()
    d_pin = torch.empty((nq, k), dtype=torch.float32).pin_memory()
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        oi, od = step_dev(q_pin.to(dev, non_blocking=True))
        ids_pin.copy_(oi, non_blocking=True); d_pin.copy_(od, non_blocking=True)
        torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t1)

    # roofline of the scan kernel: tensor pipe, 2 D operations per (vertex, query) pair of this rank's shard
    pairs = float(m) * nq
    scan_ms = float(np.mean(step_ms))
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except (OSError, ValueError):
        pass
    bf16 = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1420.0)))
    achieved = pairs * 2 * D / (scan_ms / 1e3) / 1e12
    form = {2: ("exhaustive_scan_tc16_kernel (tcgen05.mma kind::f16, M=128 N=256 K=16 x 9: 128 code dimensions + 16 threshold columns, "
                "f32 accumulators in TMEM)", "f16 x f16 -> f32 (tcgen05 kind::f16)", bf16, "MEASURED_PEAKS bf16_tflops_sustained (f16 runs at the bf16 rate; "
                "the kernel is timed inside a step)", "TFLOP/s"),
            1: ("exhaustive_scan_tc_kernel (tcgen05.mma kind::i8, M=128 N=256 K=32)", "u8 x u8 -> s32 (tcgen05 kind::i8) + f32", 2.0 * bf16,
                "2 x MEASURED_PEAKS bf16 sustained (int8 dense rate is twice bf16)", "TOP/s"),
            0: ("exhaustive_scan_kernel (popcount form)", "u32 popcount sums + f32", 2.0 * bf16, "2 x MEASURED_PEAKS bf16 sustained (int8 dense rate)", "TOP/s")}[args.scan_form]
    pieces = sharding.scan_pieces(m, world, n, args.c4_prefix, args.c4_growth) if world > 1 else [(0, m)]
    line = {"metric": "QPS (exhaustive batched scan, queries/s over the whole database)", "value": value, "unit": "queries/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": form[1], "data": "synthetic",
            "config": {"workload": f"c4: {n}x{dim} synthetic iid N(0,1), 1-bit RaBitQ codes, exhaustive scan, {nq} queries, k={k}, k'={kp}",
                       "parallelism": f"database sharded x{world}: {len(pieces)} scan pieces per shard, thresholds exchanged by NCCL all-reduce(min) between "
                                      f"them, all-gather of k' candidates (key + exact distance), merge kernel; result = one scan of the whole database",
                       "l2": "the scan streams codes + 256-query operands; candidate lists live in L2 by design"},
            "e2e": {"value": nq * args.steps / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": nq * dim * 4, "d2h_bytes_per_step": nq * k * 12},
            "gpu_launches": (9 if args.scan_form == 2 else 5) * args.steps if world == 1 else (6 * len(pieces) + 2) * args.steps,
            "roofline": {"bound": "tensor", "kernel": form[0], "achieved": achieved, "peak": form[2], "peak_source": form[3], "unit": form[4],
                         "frac": achieved / form[2], "traffic": None, "pairs_per_s_per_gpu": pairs / (scan_ms / 1e3),
                         "algorithmic_ops_per_pair": 2 * D,
                         "note": "whole step timed (K1 + threshold prefix + scan + select / re-rank, and for N > 1 the exchanges and the merge); the scan "
                                 "kernel is ~90% of it at N = 1 (DESIGN.md section 4, K5)"},
            "clocks": clocks.summary(), "step_ms": [round(x, 3) for x in step_ms]}
    if phases:
        line["phases_rank0"] = phases
    if rank == 0 and world == 1 and not args.no_recall:
        # recall@10 of (1-bit estimate -> k' candidates -> exact re-rank) against brute force, per re-rank depth k': the
        # operating point is a choice of k' (the scan's tensor-core forms reach k' = 256; beyond, the popcount form)
        ns = min(nq, 500)
        qs = q_dev[:ns]
        bd = torch.full((ns, k), float("inf"), device=dev); bi = torch.zeros((ns, k), dtype=torch.int64, device=dev)
        for s_ in range(0, m, chunk):
            bb = base[s_:s_ + chunk]
            dd = (qs * qs).sum(1, keepdim=True) - 2.0 * (qs @ bb.T) + (bb * bb).sum(1)[None, :]
            td, ti = torch.topk(dd, k, dim=1, largest=False)
            cd, ci = torch.cat([bd, td], 1), torch.cat([bi, ti + s_], 1)
            bd, sel = torch.topk(cd, k, dim=1, largest=False)
            bi = torch.gather(ci, 1, sel)
        gt = bi.cpu().numpy()
        line["recall_at_10"] = recall_at_k(out_i[:ns].cpu().numpy(), gt)
        curve = []
        for kp_ in sorted({100, 256, 1024} | {kp}):
            oi, _ = hooks.exhaustive_search(ix, qs, k, kp_)
            entry = {"kprime": kp_, "recall_at_10": recall_at_k(oi.cpu().numpy(), gt)}
            if kp_ != kp and kp_ <= 256:
                ms_, _, _, _ = timed(kp_, 2, 1)
                entry["qps"] = nq * 2 / (ms_ / 1e3)
            elif kp_ == kp:
                entry["qps"] = value
            curve.append(entry)
            log(f"[bench] c4: k'={kp_}: recall@{k} = {entry['recall_at_10']:.4f}" + (f" at {entry['qps']:.0f} QPS" if "qps" in entry else ""))
        line["recall_by_kprime"] = curve
    del ix, base
    torch.cuda.empty_cache()
    return line


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c4"],
                    help="c2: BASELINE configs[1], graph search (the default and the headline); c4: configs[3], exhaustive scan")
    ap.add_argument("--kprime", type=int, default=256, help="c4: re-rank depth (256 = the most the tensor-core scan forms keep per query)")
    ap.add_argument("--c4-n", type=int, default=10_000_000, help="database size of the c4 sub-run of the default workload")
    ap.add_argument("--c4-prefix", type=int, default=65536, help="c4, N > 1: vertices (over all shards) scanned before the first threshold exchange")
    ap.add_argument("--c4-growth", type=int, default=4, help="c4, N > 1: each later scan piece covers growth - 1 times what has been scanned")
    ap.add_argument("--c4-pieces", action="store_true", help="c4, N = 1: scan in pieces like a shard does (same result)")
    ap.add_argument("--no-c4", action="store_true", help="default workload: skip the c4 (exhaustive scan, DB-sharded) sub-run")
    ap.add_argument("--scan-form", type=int, default=2, choices=[0, 1, 2],
                    help="c4: 2 = tcgen05 kind::f16 scan with the candidate screen folded into the contraction (default), "
                         "1 = tcgen05 kind::i8 scan + float screen, 0 = popcount scan; results are identical")
    ap.add_argument("--nvec", "--n", dest="n", type=int, default=None)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--bits", type=int, default=4)
    ap.add_argument("--nq", type=int, default=10_000)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--clusters", type=int, default=0)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work for the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-recall", action="store_true")
    ap.add_argument("--no-stream", action="store_true", help="skip the K2 FastScan streaming micro-benchmark")
    ap.add_argument("--no-gate", action="store_true", help="skip the recall@10 >= 0.95 operating point")
    ap.add_argument("--gate", type=float, default=0.95)
    ap.add_argument("--gate-bits", default="4,2", help="code widths swept live for the recall gate on the n x dim shape, in this order (each needs its own "
                                                     "index: ~105 s of build per width at 1M); 1-bit at 1M runs at ~300 QPS and is reported from the recorded "
                                                     "sweep profiles/recall_sweep_1m_r02.json unless asked for here")
    ap.add_argument("--time-budget", type=float, default=240.0, help="seconds of wall time after which optional extras (further gate widths) are skipped")
    ap.add_argument("--gate-ks", default="20,40,80,160", help="k_search values of the sweep")
    ap.add_argument("--inflight", type=int, default=2, help="batches in flight (1 = each step waits for the previous one)")
    ap.add_argument("--opt", action="append", default=[], help="library tuning option name=value (does not change results)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.n is None:
        args.n = 10_000_000 if args.workload == "c4" else 1_000_000
    if args.workload == "c4":
        if args.impl == "reference":
            if rank == 0:
                print(json.dumps({"impl": "reference", "unavailable": "the reference has no exhaustive-scan mode (SURVEY F9)"}))
            return
        import torch

        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the query path has no CPU fallback")
        torch.cuda.set_device(local)
        if world > 1:
            import torch.distributed as dist

            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        line = run_c4(args, rank, world, local)
        if rank == 0:
            print(json.dumps(line))
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    metric = "QPS (search_batch queries/s; recall@10 of the reference on the same index reported beside it)"

    # ---------------- reference arm: rank 0 alone, CPU only --------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return
        path, src = obtain_index(args, 0, 1)
        q = make_queries(args, 0)
        r = time_reference(args, path, q, budget_s=90.0, steps=args.steps, warmup=args.warmup)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/cphnsw/_core*.so (the compiled reference) is not in the tree"}))
            return
        line = {"impl": "reference", "metric": metric, "value": r["value"], "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8 LUT sums + f32", "data": "synthetic", "config": run_config(args, path), "index_origin": src,
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ---------------- this repo's CUDA path ------------------------------------------------------------
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the query path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist

        from datetime import timedelta

        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=timedelta(minutes=8))   # (rank 0 may build the index: ~2 min)
    import cphnsw_b200

    path, src = obtain_index(args, rank, world)
    ix = cphnsw_b200.CPIndex(args.dim, args.bits, device=local)
    t = time.time()
    ix.load(str(path))
    info = ix.info()
    log(f"[bench] rank {rank}: index on cuda:{local} in {time.time() - t:.1f} s, {info['device_bytes'] / 2**30:.2f} GiB")
    for o in args.opt:
        name, val = o.split("=")
        ix.set_option(name, int(val))
    q = make_queries(args, rank)
    q_dev = torch.from_numpy(q).cuda()
    lib, h = ix._lib, ix.handle
    inflight = max(1, min(2, args.inflight))

    def check(rc):
        if rc:
            raise RuntimeError(lib.cphnsw_b200_last_error(h).decode())

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t_ = torch.tensor([x], dtype=torch.float64, device="cuda")
        torch.distributed.all_reduce(t_, op=torch.distributed.ReduceOp.MAX)
        return float(t_.item())

    # -- one untimed pass with the counters on (expansions etc. for the roofline), then the fast path ---
    ix.set_option("collect_stats", 1)
    ix.search_batch(q_dev, args.k)
    st = ix.last_stats()
    ix.set_option("collect_stats", 0)

    # -- value: device-resident queries and results through the C-ABI device entry point.  A step is one search_batch of
    #    the whole query batch; consecutive steps go to alternating CUDA streams (`inflight` = 2), so the drain of one
    #    batch's persistent grid overlaps the start of the next -- every step still runs the whole path on the whole batch.
    #    Timed with CUDA events: the start event precedes the first step on both streams, the end event follows the last
    #    step of both.
    ids_dev = [torch.empty((args.nq, args.k), dtype=torch.int64, device="cuda") for _ in range(inflight)]
    dists_dev = [torch.empty((args.nq, args.k), dtype=torch.float32, device="cuda") for _ in range(inflight)]
    streams = [torch.cuda.Stream() for _ in range(inflight)]

    def dev_step(i):
        s_ = i % inflight
        check(lib.cphnsw_b200_search_batch_device(h, q_dev.data_ptr(), args.nq, args.k, ids_dev[s_].data_ptr(), dists_dev[s_].data_ptr(),
                                                  streams[s_].cuda_stream))

    def timed_device_region(steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        joins = [torch.cuda.Event() for _ in streams]
        ev0.record(streams[0])
        for s_ in streams[1:]:
            s_.wait_event(ev0)
        for i in range(steps):
            dev_step(i)
        for s_, j in zip(streams[1:], joins[1:]):
            j.record(s_)
            streams[0].wait_event(j)
        ev1.record(streams[0])
        return ev0, ev1

    with ClockSampler(local, 0.01) as clocks:      # started before the warm-up so the GPU does not idle (and down-clock) before step 1
        for i in range(args.warmup):
            dev_step(i)
        barrier()
        t_begin = time.time()
        ev0, ev1 = timed_device_region(args.steps)
        barrier()
        clocks.window(t_begin, time.time())
    check(lib.cphnsw_b200_synchronize(h))
    dev_ms = max_over_ranks(ev0.elapsed_time(ev1))
    retries = ix.last_stats()["overflow_retries"]
    value = world * args.nq * args.steps / (dev_ms / 1e3)
    # the same with one batch at a time (each step waits for the previous one): the per-launch kernel time
    kernel_ms, prep_ms = [], []
    sync_ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    barrier()
    sync_ev[0].record(streams[0])
    for i in range(args.steps):
        check(lib.cphnsw_b200_search_batch_device(h, q_dev.data_ptr(), args.nq, args.k, ids_dev[0].data_ptr(), dists_dev[0].data_ptr(),
                                                  streams[0].cuda_stream))
        tm = ix.last_timings()          # waits for that step
        kernel_ms.append(tm["search_ms"]); prep_ms.append(tm["prep_ms"])
    sync_ev[1].record(streams[0])
    barrier()
    sync_ms = max_over_ranks(sync_ev[0].elapsed_time(sync_ev[1]))

    # -- e2e: the public host-buffer calls, page-locked host memory in and out, every step copies its queries to the
    #    device and its results back.  Two batches in flight (submit / wait); the one-at-a-time figure beside it.
    q_pin = [torch.from_numpy(q).pin_memory() for _ in range(inflight)]
    ids_pin = [torch.empty((args.nq, args.k), dtype=torch.int64).pin_memory() for _ in range(inflight)]
    dist_pin = [torch.empty((args.nq, args.k), dtype=torch.float32).pin_memory() for _ in range(inflight)]
    import ctypes as C

    def e2e_region(steps):
        tickets = []
        for i in range(steps):
            s_ = i % inflight
            if len(tickets) >= inflight:
                check(lib.cphnsw_b200_search_batch_wait(h, tickets.pop(0)))      # this buffer set's previous batch is home
            tk = C.c_uint64(0)
            check(lib.cphnsw_b200_search_batch_submit(h, q_pin[s_].data_ptr(), args.nq, args.k, ids_pin[s_].data_ptr(), dist_pin[s_].data_ptr(), C.byref(tk)))
            tickets.append(tk.value)
        for tk in tickets:
            check(lib.cphnsw_b200_search_batch_wait(h, tk))

    e2e_region(args.warmup)
    barrier()
    t0 = time.perf_counter()
    e2e_region(args.steps)
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * args.nq * args.steps / e2e_s
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        check(lib.cphnsw_b200_search_batch(h, q_pin[0].data_ptr(), args.nq, args.k, ids_pin[0].data_ptr(), dist_pin[0].data_ptr()))
    e2e_sync_s = max_over_ranks(time.perf_counter() - t0)
    for j in range(min(inflight, args.steps)):      # every buffer set that was written: host-buffer and device-buffer paths, both lanes
        assert np.array_equal(ids_pin[j].numpy(), ids_dev[j].cpu().numpy()), "host-buffer and device-buffer results disagree"

    # -- roofline of the dominant kernel (K3 search): algorithmic bytes / its duration.  With two batches in flight the
    #    launches overlap, so the duration charged to one launch is (timed region) / steps -- K1 and the launch gaps included.
    D, B = info["D"], info["bits"]
    block_bytes = 4 * D * B + 32 * 12 + 32 * 2 * (2 if B > 1 else 1) + 32 * 4 + 4
    vec_bytes = 4 * D + 4
    ref_exact_calls = st["nn_pushes"] + args.nq        # every exact_l2 of the reference feeds nn.push, plus the entry point
    algo_bytes = st["expansions"] * block_bytes + ref_exact_calls * vec_bytes
    k_ms = dev_ms / args.steps
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = algo_bytes / (k_ms / 1e3) / 1e9
    # DRAM traffic of this launch: bytes per expansion measured by ncu (profiles/search_kernel_traffic.json names the
    # capture) times this run's expansions -- the capture itself is never taken inside a timed run
    traffic, traffic_src = None, None
    tfile = ROOT / "profiles" / "search_kernel_traffic.json"
    if tfile.exists():
        try:
            tj = json.loads(tfile.read_text())
            if tj.get("shape") == [D, B] and tj.get("dram_bytes_per_expansion"):
                traffic = float(tj["dram_bytes_per_expansion"]) * st["expansions"]
                traffic_src = f"{tj['dram_bytes_per_expansion']:.0f} B per expansion ({tj.get('source', 'ncu')}) x this run's expansions"
        except (ValueError, KeyError):
            pass
    roofline = {"bound": "hbm", "kernel": f"search_kernel<{B}> (K3: descent + beam search + fused exact-L2 rerank)", "achieved": achieved, "peak": peak,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6.65 TB/s", "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": k_ms,
                "kernel_ms_how": f"timed region / steps with {inflight} batches in flight (launches overlap); one batch at a time the kernel "
                                 f"alone takes kernel_ms_one_at_a_time",
                "kernel_ms_one_at_a_time": float(np.mean(kernel_ms)), "frac_one_at_a_time": algo_bytes / (float(np.mean(kernel_ms)) / 1e3) / 1e9 / peak,
                "prep_kernel_ms": float(np.mean(prep_ms)),
                "expansions_per_query": st["expansions"] / args.nq, "bytes_per_expansion": block_bytes + vec_bytes,
                "frac_of_nominal_8tbs": achieved / 8000.0}

    line = {"metric": metric, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 popcount sums + f32", "data": "synthetic", "config": run_config(args, path), "index_origin": src,
            "setup": {"parallelism": f"query-sharded x{world}, index replicated", "batches_in_flight": inflight,
                      "l2": f"no flush needed: random reads over a {info['device_bytes'] / 2**30:.1f} GiB index (>> 126 MB L2)",
                      "overflow_reruns_last_step": retries},
            "value_one_batch_at_a_time": world * args.nq * args.steps / (sync_ms / 1e3),
            "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": args.nq * args.dim * 4, "d2h_bytes_per_step": args.nq * args.k * 12,
                    "ms_per_step": 1e3 * e2e_s / args.steps, "how": f"cphnsw_b200_search_batch_submit / _wait, {inflight} batches in flight, page-locked host buffers",
                    "value_one_batch_at_a_time": world * args.nq * args.steps / e2e_sync_s},
            "gpu_launches": st["kernel_launches"] * args.steps, "roofline": roofline, "clocks": clocks.summary(),
            "kernel_ms_one_at_a_time": [round(x, 3) for x in kernel_ms],
            "search_stats_per_query": {k: v / args.nq for k, v in st.items() if k not in ("max_beam", "overflow_retries", "kernel_launches")} | {"max_beam": st["max_beam"]}}
    ids_dev, dists_dev = ids_dev[0], dists_dev[0]

    # -- K2 alone: the FastScan estimator streaming every neighbour block of the index once (one query) ---
    if rank == 0 and not args.no_stream:
        from cphnsw_b200 import hooks

        prep = hooks.prepare_queries(ix, q_dev[:1])
        nblk = info["n"]
        dqp = torch.full((nblk,), 200.0, dtype=torch.float32, device="cuda")
        outs = {n_: torch.empty((nblk, 32), dtype=torch.float32, device="cuda") for n_ in ("est", "lower")}
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        times = []
        for it in range(args.warmup + args.steps):
            ev[0].record()
            hooks.fastscan_blocks(ix, prep["uplanes"], prep["coeffs"], dqp, first_vertex=0, nblocks=nblk, want=("est", "lower"), out=outs)
            ev[1].record()
            torch.cuda.synchronize()
            if it >= args.warmup:
                times.append(ev[0].elapsed_time(ev[1]))
        del outs
        fs_bytes = 4 * D * B + 384 + 64 * (2 if B > 1 else 1)            # per 32-code block, ids excluded
        log(f"[bench] K2 stream times (ms): {[round(t, 3) for t in times]}")
        fs_ms = float(np.mean(times))
        fs_gbs = nblk * fs_bytes / (fs_ms / 1e3) / 1e9
        line["fastscan_stream"] = {"kernel": f"fastscan_blocks_kernel<{B}> (K2, TMA-pipelined stream of all {nblk} blocks, one query, est+lower written)",
                                   "achieved": fs_gbs, "unit": "GB/s", "peak": peak, "frac": fs_gbs / peak, "frac_of_nominal_8tbs": fs_gbs / 8000.0, "ms": fs_ms,
                                   "algorithmic_bytes_per_block": fs_bytes, "blocks_per_s": nblk / (fs_ms / 1e3), "codes_per_s": 32 * nblk / (fs_ms / 1e3),
                                   # SURVEY 8(d)'s figure counts what a block READS; this stand-alone kernel also writes est + lower (2 x 128 B per block)
                                   "output_bytes_per_block": 256, "achieved_with_output": nblk * (fs_bytes + 256) / (fs_ms / 1e3) / 1e9,
                                   "frac_with_output": nblk * (fs_bytes + 256) / (fs_ms / 1e3) / 1e9 / peak}

    # -- BASELINE config 4 beside it: the exhaustive batched scan over 10M x 96 1-bit codes, database sharded over the ranks
    #    (strong scaling; NCCL threshold exchange + all-gather + merge kernel).  All ranks take part.
    if not args.no_c4:
        import copy

        ca = copy.copy(args)
        ca.n, ca.steps, ca.warmup = args.c4_n, max(2, min(args.steps, 5)), max(1, min(args.warmup, 2))
        try:
            c4 = run_c4(ca, rank, world, local)
            line["c4"] = {k_: c4[k_] for k_ in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "config", "e2e",
                                                "gpu_launches", "roofline", "clocks", "step_ms", "recall_at_10", "recall_by_kprime") if k_ in c4}
        except Exception as e:  # noqa: BLE001 - a secondary figure must not take the bench line down
            log(f"[bench] rank {rank}: c4 sub-run skipped: {type(e).__name__}: {e}")

    if rank == 0:
        ids_np = ids_dev.cpu().numpy()
        if not args.no_recall:
            gt = ground_truth(path, q, args.k, torch.device("cuda", local))
            line["recall_at_10"] = recall_at_k(ids_np, gt)
            line["unique_ids_per_row"] = float(np.mean([len(set(r.tolist())) for r in ids_np]))
        if world == 1 and not args.no_cpu_baseline:
            r = time_reference(args, path, q, budget_s=args.cpu_budget, steps=1, warmup=0)
            if r is not None:
                line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
                line["cpu_baseline"]["cpu_model"] = cpu_model()
                one = reference_one_thread(args, path, args.k)
                if one:
                    line["cpu_baseline"]["one_thread"] = {"value": one["value"], "unit": "queries/s", "sample": f"{one['queries']} queries, OMP_NUM_THREADS=1"}
                same_i, same_d = same_results(r["ids"], r["dists"], ids_np[:r["m"]], dists_dev[:r["m"]].cpu().numpy())
                line["cpu_baseline"]["ids_identical_to_gpu"] = same_i
                line["cpu_baseline"]["distance_bits_identical_to_gpu"] = same_d
                if not args.no_recall:
                    line["cpu_baseline"]["recall_at_10"] = recall_at_k(r["ids"], gt[:r["m"]])
        # -- the operating point the metric names: recall@10 >= 0.95 on n x 128.  Swept over the code widths in
        #    --gate-bits (each its own reference-built index) x k_search; the first (bits, k_search) that clears the gate is
        #    `recall_gate`, with both arms' QPS there; every curve is kept either way, so "not reachable" is a measurement.
        if world == 1 and not args.no_recall and not args.no_gate:
            ks_list = [int(x) for x in args.gate_ks.split(",")]
            sweeps, gate = [], None
            for gb in [int(x) for x in args.gate_bits.split(",") if x]:
                need_s = 0.0 if gb == args.bits else 130.0 * args.n / 1e6      # index build + sweep of another width
                if time.time() - T_START + need_s > args.time_budget:
                    log(f"[bench] recall sweep at {gb} bits skipped: time budget ({args.time_budget:.0f} s)")
                    continue
                try:
                    sw = recall_sweep(args, local, torch, cphnsw_b200, args.n, gb, ks_list, ix=ix if gb == args.bits else None)
                except Exception as e:  # noqa: BLE001 - a secondary figure must not take the bench line down
                    log(f"[bench] recall sweep at {gb} bits skipped: {e}")
                    continue
                sweeps.append(sw)
                if sw["reached"]:
                    gate = sw
                    break
            line["recall_gate"] = dict(gate) if gate else {"reached": False, "gate": args.gate,
                                                           "note": "no (bits, k_search) tried clears the gate on this data; both arms return identical ids, "
                                                                   "so this is the reference's own search quality (SURVEY H5)"}
            line["recall_gate"]["sweeps"] = [{k_: v for k_, v in sw.items() if k_ in ("bits", "n", "curve", "reached", "index_file")} for sw in sweeps]
            rec = ROOT / "profiles" / "recall_sweep_1m_r02.json"
            if rec.exists() and args.n == 1_000_000 and args.dim == 128 and not args.clusters:
                try:      # widths not swept live in this run, as recorded by the builder with the same code path (file says how)
                    rj = json.loads(rec.read_text())
                    live = {sw["bits"] for sw in sweeps}
                    line["recall_gate"]["sweeps_recorded"] = {"source": "profiles/recall_sweep_1m_r02.json", "how": rj.get("how"),
                                                              "sweeps": [sw for sw in rj.get("sweeps", []) if sw.get("bits") not in live]}
                except ValueError:
                    pass
            del ix
            if not gate or gate["n"] != 100_000:
                try:      # BASELINE configs[0]: the shape on which the reference itself is known to clear the gate
                    line["recall_gate_100k_1bit"] = recall_sweep(args, local, torch, cphnsw_b200, 100_000, 1, (10, 20, 30, 40))
                except Exception as e:  # noqa: BLE001
                    log(f"[bench] gate config skipped: {e}")
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
