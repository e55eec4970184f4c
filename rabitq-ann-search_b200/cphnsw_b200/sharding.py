"""Multi-GPU partitioning of the query path: one process per GPU, `torch.distributed` for plumbing.

Two modes:

* graph search -- queries are independent (the reference fans them out over OpenMP threads,
  src/bindings.cpp:196-200): every rank holds the whole index, searches a contiguous slice of the
  queries, and the slices are concatenated.  No data-path collective; the optional all-gather only
  serves callers that want the full result on every rank.
* exhaustive scan -- the database is split into contiguous internal-id ranges, every rank scans its
  range for all queries and keeps a local top-k; one all-gather of [nq, k] (id, distance) pairs and a
  local k-way merge ordered by (distance, id) give the global top-k.  Exact: the global top-k is a
  subset of the union of the per-shard top-k lists.

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


def exhaustive_search_db_sharded(local_scan, n: int, queries: np.ndarray, k: int, kprime: int, group=None):
    """Exhaustive scan with the database split across ranks.

    local_scan(queries, k, kprime, id_begin, id_end) -> (ids [nq,k] global internal ids, dists [nq,k]) is the
    rank's own scan of its id range.  One all-gather (NCCL over NVLink when the group is nccl) of the
    per-shard top-k, then a local merge; every rank returns the global result."""
    import torch

    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    b, e = db_shard(n, rank, world)
    ids, dists = local_scan(queries, k, kprime, b, e)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    ti = torch.as_tensor(ids, device=dev).to(torch.int64).contiguous()
    td = torch.as_tensor(dists, device=dev).to(torch.float32).contiguous()
    all_i = [torch.empty_like(ti) for _ in range(world)]
    all_d = [torch.empty_like(td) for _ in range(world)]
    dist.all_gather(all_i, ti, group=group)
    dist.all_gather(all_d, td, group=group)
    return merge_topk(torch.stack(all_i).cpu().numpy(), torch.stack(all_d).cpu().numpy(), k)
