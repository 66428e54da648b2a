"""Multi-GPU plumbing of the PPP scan.

Every (query, target) pair is independent, so the path shards with no data-path collective: the TARGET axis is split
into contiguous shards, one process per GPU holds its shard resident and receives all queries.  The only exchange
step is the optional global top-k merge: an all-gather of the per-rank per-query top-k lists (k x 32 B per query per
rank) followed by a k-way merge with the same (score, target index) order the device merge kernel uses.
Works with any torch.distributed backend (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np

from .capi import HIT_DTYPE


def shard_range(n_targets: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous target shard [t0, t1) of `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_targets, world)
    t0 = rank * base + min(rank, rem)
    return t0, t0 + base + (1 if rank < rem else 0)


def merge_topk(parts: list[np.ndarray], k: int, n_queries: int) -> np.ndarray:
    """k-way merge of per-shard top-k hit lists (target indices already global): ascending score, ties by target
    index -- the order of bin/ppp.pl:473-474's sort plus a deterministic tie-break."""
    allh = np.concatenate(parts) if parts else np.empty(0, dtype=HIT_DTYPE)
    if allh.size == 0:
        return allh
    order = np.lexsort((allh["t"], allh["score"], allh["q"]))
    allh = allh[order]
    first = np.searchsorted(allh["q"], np.arange(n_queries), side="left")
    rank_in_q = np.arange(allh.size) - first[allh["q"]]
    return allh[rank_in_q < k]


def all_gather_hits(hits: np.ndarray, t_offset: int, device=None):
    """All-gather variable-length hit lists over the default process group; returns the list of per-rank arrays with
    target indices shifted to global numbering (rank r adds its shard offset before sending)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    h = hits.copy()
    h["t"] += np.uint32(t_offset)
    loc = torch.from_numpy(h.view(np.uint8).reshape(-1).copy())
    if device is not None:
        loc = loc.to(device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=loc.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([loc.numel()], dtype=torch.int64, device=loc.device))
    mx = int(max(int(s.item()) for s in sizes))
    buf = torch.zeros(mx, dtype=torch.uint8, device=loc.device)
    buf[: loc.numel()] = loc
    allb = [torch.empty(mx, dtype=torch.uint8, device=loc.device) for _ in range(world)]
    dist.all_gather(allb, buf)
    return [b[: int(s.item())].cpu().numpy().view(HIT_DTYPE).copy() for b, s in zip(allb, sizes)]


def global_topk(hits: np.ndarray, t_offset: int, k: int, n_queries: int, device=None) -> np.ndarray:
    """Per-rank top-k lists -> the global top-k of every query (identical on all ranks)."""
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        h = hits.copy()
        h["t"] += np.uint32(t_offset)
        return merge_topk([h], k, n_queries)
    return merge_topk(all_gather_hits(hits, t_offset, device), k, n_queries)


class _DevBuf:
    """Expose a raw device allocation of the library to torch through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


def global_topk_device(ctx, t_offset_of_rank, k: int, n_queries: int, device, out=None, cnt=None):
    """Global top-k over the ranks' target shards without a host round trip: NCCL all-gather of the DEVICE-resident
    per-rank lists (ppp_topk_device), then the library's merge kernel (ppp_merge_topk_device).  Returns (hits, counts)
    on the host, identical on every rank.  t_offset_of_rank(r) = first global target index of rank r's shard."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    lp, cp, kk, nq = ctx.topk_device()
    assert kk == k and nq == n_queries
    lists = torch.as_tensor(_DevBuf(lp, nq * k * HIT_DTYPE.itemsize), device=device)
    counts = torch.as_tensor(_DevBuf(cp, nq * 4), device=device)
    g_lists = torch.empty(world * lists.numel(), dtype=torch.uint8, device=device)
    g_counts = torch.empty(world * counts.numel(), dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(g_lists, lists)
    dist.all_gather_into_tensor(g_counts, counts)
    torch.cuda.synchronize(device)
    offs = [t_offset_of_rank(r) for r in range(world)]
    return ctx.merge_topk_device(g_lists.data_ptr(), g_counts.data_ptr(), world, offs, n_queries, k, out=out, cnt=cnt)
