"""Multi-GPU plumbing of the PPP scan.

Every (query, target) pair is independent, so the path shards with no data-path collective: the TARGET axis is split
into contiguous shards, one process per GPU holds its shard resident and receives all queries.  The only exchange
step is the optional global top-k merge.  Its cheap form (SlicedTopkMerge) splits the QUERY axis for the merge: rank r
receives every rank's lists of its own slice of the queries (one all-to-all: k x 32 B per query in total per rank instead
of N times that), merges 1/N of the queries on its GPU and holds that slice of the global result.  The replicated form
(all-gather + full merge on every rank) is kept for callers that want the whole result everywhere.
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


def balanced_query_shards(P: np.ndarray, Y: np.ndarray, G: int, world: int) -> list[np.ndarray]:
    """Query-axis partition for `world` GPUs: sorted index arrays, one per rank, covering every query once.

    A query's cost in the scan is set by its presence plane (an all-ones plane takes `number` from the target's depth prefix: 3
    POPC per 128 genomes; any other plane counts T & P as well: 6) and, to a lesser degree, by |Y| (staircase length, pairs that
    reach the full evaluation).  Contiguous slices of a random query set differ by a few percent in their number of general
    planes, and the step time of an N-GPU job is the slowest rank's.  So the queries of each class are dealt over the ranks
    in descending |Y| order, serpentine (0..N-1, N-1..0, ...) -- the second class continues where the first stopped, so rank sizes
    differ by at most one query.  The reference has no counterpart (bin/ppp.pl:895-922 forks one client per 50 targets and lets the OS
    balance them)."""
    wpr = P.shape[1]
    full = np.full(wpr, 0xFFFFFFFF, dtype=np.uint32)
    if G % 32:
        full[G // 32] = (1 << (G % 32)) - 1
    full[(G + 31) // 32:] = 0
    all_ones = (P == full[None, :]).all(axis=1)
    ny = np.zeros(P.shape[0], dtype=np.int64)
    for w in range(wpr):  # |Y| per query without unpacking the planes
        ny += _popcount32(Y[:, w])
    shards = [[] for _ in range(world)]
    i = 0
    for cls in (np.nonzero(~all_ones)[0], np.nonzero(all_ones)[0]):
        for q in cls[np.argsort(-ny[cls], kind="stable")]:
            rnd, j = divmod(i, world)
            shards[j if rnd % 2 == 0 else world - 1 - j].append(int(q))  # serpentine: 0..N-1, N-1..0, ... evens out the |Y| sums
            i += 1
    return [np.array(sorted(sh), dtype=np.int64) for sh in shards]


def _popcount32(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint32)
    x = x - ((x >> 1) & 0x55555555)
    x = (x & 0x33333333) + ((x >> 2) & 0x33333333)
    x = (x + (x >> 4)) & 0x0F0F0F0F
    return ((x * 0x01010101) >> 24).astype(np.int64) & 0xFF


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


def global_topk_sliced_host(lists: np.ndarray, counts: np.ndarray, t_offset: int, k: int, n_queries: int):
    """Host-memory form of the sliced merge (any backend; the gloo tests run it on CPU): `lists` = this rank's padded
    [n_queries, k] top-k lists with LOCAL target indices, `counts` [n_queries].  Rank r receives every rank's lists of its
    query slice (all-to-all), shifts them to global target numbering and k-way merges them.  Returns (q0, hits) -- the
    merged lists of queries [q0, q1) of this rank, grouped by query."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(), dist.get_rank()
    bounds = [shard_range(n_queries, r, world) for r in range(world)]
    q0, q1 = bounds[rank]
    n_slice, rec = q1 - q0, HIT_DTYPE.itemsize
    send_rows = [b - a for a, b in bounds]
    h = np.ascontiguousarray(lists).reshape(n_queries, k).copy()
    h["t"] += np.uint32(t_offset)
    src = torch.from_numpy(h.view(np.uint8).reshape(-1))
    cnt = torch.from_numpy(np.ascontiguousarray(counts, dtype=np.uint32).view(np.uint8).reshape(-1).copy())
    g_lists = torch.empty(world * n_slice * k * rec, dtype=torch.uint8)
    g_counts = torch.empty(world * n_slice * 4, dtype=torch.uint8)
    dist.all_to_all_single(g_lists, src, output_split_sizes=[n_slice * k * rec] * world, input_split_sizes=[r * k * rec for r in send_rows])
    dist.all_to_all_single(g_counts, cnt, output_split_sizes=[n_slice * 4] * world, input_split_sizes=[r * 4 for r in send_rows])
    gl = g_lists.numpy().view(HIT_DTYPE).reshape(world, n_slice, k)
    gc = g_counts.numpy().view(np.uint32).reshape(world, n_slice)
    parts = [gl[r, j, : gc[r, j]] for r in range(world) for j in range(n_slice)]
    merged = merge_topk([x for x in parts if x.size], k, n_queries)
    return q0, merged


class SlicedTopkMerge:
    """Global top-k of a target-sharded run, merged by query slices.

    After ctx.run(top-k) on every rank: rank r receives, from every rank s, the DEVICE-resident lists of the queries
    [q0_r, q1_r) (shard_range over the query axis) -- one NCCL all-to-all straight out of the library's buffers
    (ppp_topk_device), 1/N of the traffic of an all-gather -- merges them with the library's merge kernel
    (ppp_merge_topk_device: 1/N of the merge work per GPU) and copies its slice of the global lists into a pinned host buffer.
    The job's result is the N slices; `gather()` assembles them on every rank for callers that want that."""

    def __init__(self, ctx, n_queries: int, k: int, world: int, rank: int, device):
        import torch

        self.ctx, self.nq, self.k, self.world, self.rank, self.device = ctx, n_queries, k, world, rank, device
        self.bounds = [shard_range(n_queries, r, world) for r in range(world)]
        self.q0, self.q1 = self.bounds[rank]
        self.n_slice = self.q1 - self.q0
        rec = HIT_DTYPE.itemsize
        self.send_rows = [b - a for a, b in self.bounds]
        self.g_lists = torch.empty(world * self.n_slice * k * rec, dtype=torch.uint8, device=device)
        self.g_counts = torch.empty(world * self.n_slice * 4, dtype=torch.uint8, device=device)
        self.out_t = torch.empty(max(1, self.n_slice * k * rec), dtype=torch.uint8).pin_memory()
        self.cnt_t = torch.empty(max(1, self.n_slice), dtype=torch.int32).pin_memory()
        self.out = self.out_t.numpy()[: self.n_slice * k * rec].view(HIT_DTYPE)
        self.cnt = self.cnt_t.numpy()[: self.n_slice].view(np.uint32)
        self.d2h_bytes = self.n_slice * k * rec + self.n_slice * 4
        self.launches_per_run = 1  # the merge kernel (the all-to-all's copy kernels are NCCL's)

    def run(self, t_offset_of_rank):
        """-> (hits [n_slice * k], counts [n_slice]) of this rank's query slice, target indices global.
        t_offset_of_rank(r) = first global target index of rank r's shard."""
        import torch
        import torch.distributed as dist

        k, rec = self.k, HIT_DTYPE.itemsize
        lp, cp, kk, nq = self.ctx.topk_device()
        assert kk == k and nq == self.nq
        lists = torch.as_tensor(_DevBuf(lp, nq * k * rec), device=self.device)
        counts = torch.as_tensor(_DevBuf(cp, nq * 4), device=self.device)
        dist.all_to_all_single(self.g_lists, lists, output_split_sizes=[self.n_slice * k * rec] * self.world,
                               input_split_sizes=[r * k * rec for r in self.send_rows])
        dist.all_to_all_single(self.g_counts, counts, output_split_sizes=[self.n_slice * 4] * self.world,
                               input_split_sizes=[r * 4 for r in self.send_rows])
        torch.cuda.current_stream(self.device).synchronize()  # the merge runs on the library's own stream
        offs = [t_offset_of_rank(r) for r in range(self.world)]
        if self.n_slice:
            self.ctx.merge_topk_device(self.g_lists.data_ptr(), self.g_counts.data_ptr(), self.world, offs, self.n_slice, k, out=self.out, cnt=self.cnt)
        return self.out, self.cnt

    def gather(self):
        """the slices of the last run() assembled on every rank: (hits [nq * k] padded lists, counts [nq])"""
        import torch
        import torch.distributed as dist

        rec = HIT_DTYPE.itemsize
        mx = max(self.send_rows)
        loc = torch.zeros(mx * self.k * rec, dtype=torch.uint8, device=self.device)
        loc[: self.n_slice * self.k * rec] = self.out_t[: self.n_slice * self.k * rec].to(self.device)
        lc = torch.zeros(mx, dtype=torch.int32, device=self.device)
        lc[: self.n_slice] = self.cnt_t[: self.n_slice].to(self.device)
        allh = torch.empty(self.world * loc.numel(), dtype=torch.uint8, device=self.device)
        allc = torch.empty(self.world * mx, dtype=torch.int32, device=self.device)
        dist.all_gather_into_tensor(allh, loc)
        dist.all_gather_into_tensor(allc, lc)
        h = allh.cpu().numpy().view(HIT_DTYPE).reshape(self.world, mx * self.k)
        c = allc.cpu().numpy().view(np.uint32).reshape(self.world, mx)
        hits = np.concatenate([h[r, : self.send_rows[r] * self.k] for r in range(self.world)])
        cnts = np.concatenate([c[r, : self.send_rows[r]] for r in range(self.world)])
        return hits, cnts


def check_global_lists(ctx, merger: SlicedTopkMerge, t_offset_of_rank, P, Y, p, T_local, G: int, k: int, rank: int, world: int, device,
                       n_check: int = 64, rel_tol: float = 1e-9):
    """Post-run self-check of a multi-GPU top-k job (bench.py, N > 1): for n_check queries (the first n_check / N of every
    rank's slice) the merged global list must equal the list ONE GPU computes over the concatenated shards.  Every rank
    all-gathers the target planes (NVLink), scores its check queries against all of them on its own GPU and compares with
    its slice of the merged result: targets, counts and cut-off indices exactly, scores within rel_tol on -ln p (a pair can
    take the series tier in one run and the exact tier in the other).  Leaves other targets/queries resident in ctx."""
    import torch
    import torch.distributed as dist

    from . import capi

    merged, counts = merger.run(t_offset_of_rank)
    merged, counts = merged.copy(), counts.copy()
    rows = torch.tensor([T_local.shape[0]], dtype=torch.int64, device=device)
    all_rows = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(all_rows, rows)
    all_rows = [int(r.item()) for r in all_rows]
    assert len(set(all_rows)) == 1, "check_global_lists expects equal shards"
    loc = torch.from_numpy(np.ascontiguousarray(T_local).view(np.int32)).to(device)  # NCCL has no uint32
    full = torch.empty((world * T_local.shape[0], T_local.shape[1]), dtype=loc.dtype, device=device)
    dist.all_gather_into_tensor(full, loc)
    full_np = full.cpu().numpy().view(np.uint32)
    del full, loc
    order = np.argsort([t_offset_of_rank(r) for r in range(world)], kind="stable")
    if list(order) != list(range(world)):  # shards are gathered in rank order; global target ids follow the offsets
        full_np = np.concatenate([full_np[r * all_rows[0]:(r + 1) * all_rows[0]] for r in order])
    m = min(max(1, n_check // world), merger.n_slice)
    qs = np.arange(merger.q0, merger.q0 + m)
    ctx.set_targets_bits(full_np, G)
    one = ctx.score(np.ascontiguousarray(P[qs]), np.ascontiguousarray(Y[qs]), np.ascontiguousarray(p[qs]), capi.make_filter(capi.FILTER_TOPK, k=k))
    base = min(t_offset_of_rank(r) for r in range(world))
    ok = True
    kk = min(k, full_np.shape[0])
    for j in range(m):
        a = merged[j * k: j * k + int(counts[j])]
        b = one[one["q"] == j]
        if a.size != b.size or a.size != kk:
            ok = False
            break
        same = np.array_equal(a["t"], b["t"] + np.uint32(base)) and all(np.array_equal(a[f], b[f]) for f in ("yes", "number", "depth"))
        with np.errstate(divide="ignore", invalid="ignore"):
            rel = np.abs(np.log(a["score"]) - np.log(b["score"])) / np.abs(np.log(b["score"]))
        same = same and bool(np.all((a["score"] == b["score"]) | (rel <= rel_tol))) and bool(np.all(a["q"] == qs[j]))
        if not same:
            ok = False
            break
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return {"queries_checked": int(m * world), "targets": int(full_np.shape[0]), "merged_lists_equal_single_gpu_lists": bool(flag.item() == 1)}
