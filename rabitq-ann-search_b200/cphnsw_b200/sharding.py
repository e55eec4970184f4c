"""Multi-GPU partitioning of the query path: one process per GPU, `torch.distributed` for plumbing.

Two modes:

* graph search -- queries are independent (the reference fans them out over OpenMP threads,
  src/bindings.cpp:196-200): every rank holds the whole index, searches a contiguous slice of the
  queries, and the slices are concatenated.  No data-path collective; the optional all-gather only
  serves callers that want the full result on every rank.
* exhaustive scan -- the database is split into contiguous internal-id ranges; every rank scans its range
  for all queries and keeps its k' best (estimate, id) candidates with their exact distances; one
  all-gather of those and a merge on every rank (the k' best estimates overall, then the k best
  distances among them) give exactly what one scan of the whole database gives.  Thresholds are
  exchanged on the way (all-reduce(min) of nq floats at a few points) so that a shard's work shrinks with
  its size instead of every shard re-discovering them.

`local_search` arguments make the plumbing testable on CPU (gloo) with any callable.
"""
from __future__ import annotations

import numpy as np

_FLT_MAX = np.finfo(np.float32).max


def query_shard(nq: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced slice [begin, end) of the queries for `rank`."""
    base, extra = divmod(nq, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def db_shard(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced internal-id range [begin, end) of the database for `rank`."""
    return query_shard(n, rank, world)


def merge_topk(ids: np.ndarray, dists: np.ndarray, k: int):
    """k-way merge of per-shard top-k lists: ids/dists [shards, nq, k'] -> [nq, k], ordered by
    (distance, id); padding entries (id < 0) sort last."""
    s, nq, kk = ids.shape
    ii = np.transpose(ids, (1, 0, 2)).reshape(nq, s * kk)
    dd = np.transpose(dists, (1, 0, 2)).reshape(nq, s * kk).astype(np.float32)
    pad = ii < 0
    key_id = np.where(pad, np.iinfo(np.int64).max, ii)
    key_d = np.where(pad, np.float32(np.inf), dd)
    order = np.lexsort((key_id, key_d), axis=1)[:, :k]
    out_i = np.take_along_axis(ii, order, 1)
    out_d = np.take_along_axis(dd, order, 1)
    if out_i.shape[1] < k:
        fill = k - out_i.shape[1]
        out_i = np.concatenate([out_i, np.full((nq, fill), -1, np.int64)], 1)
        out_d = np.concatenate([out_d, np.full((nq, fill), _FLT_MAX, np.float32)], 1)
    out_d = np.where(out_i < 0, _FLT_MAX, out_d).astype(np.float32)
    return out_i.astype(np.int64), out_d


def merge_topk_device(all_ids, all_dists, k: int):
    """merge_topk on the device: torch tensors [shards, nq, k'] -> ([nq, k] int64, [nq, k] float32), the same
    (distance, id) order.  Distances are >= 0, so their bit patterns order like the floats; ids are < 2^32."""
    import torch

    s, nq, kk = all_ids.shape
    ii = all_ids.permute(1, 0, 2).reshape(nq, s * kk)
    dd = all_dists.permute(1, 0, 2).reshape(nq, s * kk).contiguous()
    key = (dd.view(torch.int32).to(torch.int64) << 32) | (ii & 0xFFFFFFFF)
    key = torch.where(ii < 0, torch.full_like(key, torch.iinfo(torch.int64).max), key)
    kk_out = min(k, s * kk)
    sel = torch.topk(key, kk_out, dim=1, largest=False, sorted=True).indices
    out_i = torch.gather(ii, 1, sel)
    out_d = torch.gather(dd, 1, sel)
    out_d = torch.where(out_i < 0, torch.full_like(out_d, float(_FLT_MAX)), out_d)
    if kk_out < k:
        out_i = torch.cat([out_i, torch.full((nq, k - kk_out), -1, dtype=out_i.dtype, device=out_i.device)], 1)
        out_d = torch.cat([out_d, torch.full((nq, k - kk_out), float(_FLT_MAX), dtype=out_d.dtype, device=out_d.device)], 1)
    return out_i, out_d


def _dist():
    import torch.distributed as dist

    return dist


def search_batch_query_sharded(local_search, queries: np.ndarray, k: int, gather: bool = True, group=None):
    """Graph search over `queries` split across the ranks of `group` (index replicated).

    local_search(q_slice, k) -> (ids, dists) is the rank's own CPIndex.search_batch.  Returns the rank's
    slice, or with gather=True the full [nq, k] result on every rank."""
    import torch

    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    nq = queries.shape[0]
    b, e = query_shard(nq, rank, world)
    ids, dists = local_search(queries[b:e], k)
    if not gather:
        return ids, dists
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    cap = max(query_shard(nq, r, world)[1] - query_shard(nq, r, world)[0] for r in range(world))
    pad_i = torch.full((cap, k), -1, dtype=torch.int64, device=dev)
    pad_d = torch.full((cap, k), float(_FLT_MAX), dtype=torch.float32, device=dev)
    pad_i[: e - b] = torch.as_tensor(ids, device=dev)
    pad_d[: e - b] = torch.as_tensor(dists, device=dev)
    all_i = [torch.empty_like(pad_i) for _ in range(world)]
    all_d = [torch.empty_like(pad_d) for _ in range(world)]
    dist.all_gather(all_i, pad_i, group=group)
    dist.all_gather(all_d, pad_d, group=group)
    out_i = np.concatenate([all_i[r][: query_shard(nq, r, world)[1] - query_shard(nq, r, world)[0]].cpu().numpy() for r in range(world)])
    out_d = np.concatenate([all_d[r][: query_shard(nq, r, world)[1] - query_shard(nq, r, world)[0]].cpu().numpy() for r in range(world)])
    return out_i, out_d


def scan_pieces(n_local: int, world: int, n_total: int | None = None, prefix: int = 65536, growth: int = 4) -> list[tuple[int, int]]:
    """Ranges [begin, end) in which a shard of n_local vertices is scanned.  The first piece is the shard's part of the
    threshold prefix (the shards' first pieces together hold `prefix` vertices, what a single scan uses to find its
    first thresholds); every later piece is `growth` - 1 times what has been scanned so far, so thresholds are exchanged
    after 1/world of the prefix, then at geometrically spaced points -- a handful of all-reduces of nq floats.
    The plan is laid out for the largest shard (ceil(n_total / world)) and clipped to n_local, so that every rank makes
    the same number of exchanges (a smaller shard's last piece may be empty)."""
    world = max(world, 1)
    n_ref = max(n_local, -(-(n_total if n_total is not None else n_local * world) // world))
    first = max(1024, -(-prefix // world))
    out, b, size = [], 0, first
    while True:
        e = min(n_ref, b + size)
        if n_ref - e < size // 2:     # do not leave a sliver
            e = n_ref
        out.append((min(b, n_local), min(e, n_local)))
        if e >= n_ref:
            return out
        size = e * (growth - 1)
        b = e


def _all_gather_stacked(t, world, group):
    """[world, *t.shape] on every rank (one NCCL all-gather into a single tensor; gloo gathers into a list)."""
    import torch

    dist = _dist()
    t = t.contiguous()
    if dist.get_backend(group) == "nccl":
        out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t, group=group)
        return out
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t, group=group)
    return torch.stack(parts)


def kth_estimate(keys, m: int):
    """float32 [nq]: the m-th smallest estimate among each row of candidate keys (int64 tensor holding estimate bits << 32 |
    id, -1 = padding); FLT_MAX where a row holds fewer than m keys.  Estimates are non-negative floats, so their bit
    patterns order like the values."""
    import torch

    est = (keys >> 32) & 0xFFFFFFFF
    est = torch.where(keys == -1, torch.full_like(est, 0x7F7FFFFF), est)
    return torch.kthvalue(est, m, dim=1).values.to(torch.int32).view(torch.float32)


def exhaustive_search_db_sharded(local_candidates, merge_candidates, n_local: int, n_total: int, id_offset: int, queries, k: int,
                                 kprime: int, group=None, prefix: int = 65536, growth: int = 4, timeline: list | None = None):
    """Exhaustive scan with the database split across the ranks of `group`; returns on every rank what ONE scan of the
    whole database returns: the kprime smallest (estimate, id) overall, exact distances, the k smallest (distance, id).

    local_candidates(queries, kprime, begin, end, id_offset, prior_keys, tau_in, want_dists) -> (keys, dists, tau) and
    merge_candidates(keys [lists, nq, kprime], dists | None, k) -> (ids, dists, tau) are hooks.exhaustive_candidates /
    hooks.merge_candidates bound to the rank's index (tests pass CPU stand-ins); tensors live wherever those put them.

    Every shard scans its range in pieces (scan_pieces).  After the first piece the first common threshold is the
    all-reduce(MAX) over the shards of each shard's ceil(kprime / world)-th smallest estimate: every shard then holds at least
    that many keys under it, the union at least kprime, so it bounds the kprime-th smallest of the union from above -- and for
    shards of one distribution it sits at the same quantile (measured at 8 ranks: 0.1 ms instead of the 1.6 ms an
    all-gather of the keys and a merge took).  After every later piece one all-reduce(min) of the per-query thresholds.  A threshold is only ever an upper bound of the final
    kprime-th estimate, so dropping pairs above it cannot change the result -- it only keeps a shard from collecting
    candidates the other shards have already ruled out.  At the end: one all-gather of every shard's kprime keys +
    exact distances, merged on every rank."""
    import torch

    dist = _dist() if group is not None or _dist().is_initialized() else None
    world = dist.get_world_size(group) if dist else 1
    pieces = scan_pieces(n_local, world, n_total, prefix, growth)

    def mark(name):      # optional device timeline (CUDA events on the current stream): [(phase name, event), ...]
        if timeline is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            timeline.append((name, ev))

    keys = dists = tau = None
    mark("start")
    for c, (b, e) in enumerate(pieces):
        last = c == len(pieces) - 1
        keys, dists, tau_local = local_candidates(queries, kprime, b, e, id_offset, keys, tau, last)
        mark(f"piece {c} [{b}, {e})")
        if last:
            break
        if world == 1:
            tau = tau_local
        elif c == 0:
            tau = kth_estimate(keys, -(-kprime // world)).contiguous()
            dist.all_reduce(tau, op=dist.ReduceOp.MAX, group=group)
            mark("all-reduce(max) of the shards' ceil(k'/world)-th estimates")
        else:
            tau = tau_local.clone()
            dist.all_reduce(tau, op=dist.ReduceOp.MIN, group=group)
            mark("all-reduce(min) of thresholds")
    if world == 1:
        out = merge_candidates(keys[None], dists[None], k)[:2]
    else:
        gk, gd = _all_gather_stacked(keys, world, group), _all_gather_stacked(dists, world, group)
        mark("all-gather of candidates")
        out = merge_candidates(gk, gd, k)[:2]
    mark("merge")
    return out


def exhaustive_search_db_sharded_device(ix, queries, k: int, kprime: int, n_local: int, n_total: int, id_offset: int, group=None,
                                        prefix: int = 65536, growth: int = 4, timeline: list | None = None):
    """exhaustive_search_db_sharded on the rank's `ix` (a cphnsw_b200.CPIndex holding the shard); CUDA tensors throughout,
    NCCL for the exchanges."""
    from . import hooks

    def cand(q, kp, b, e, off, prior, tau, want):
        return hooks.exhaustive_candidates(ix, q, kp, b, e, off, prior, tau, want)

    def merge(keys, dists, kk):
        return hooks.merge_candidates(ix, keys, dists, kk)

    return exhaustive_search_db_sharded(cand, merge, n_local, n_total, id_offset, queries, k, kprime, group, prefix, growth, timeline)
