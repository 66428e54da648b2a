"""CPU tests of the multi-GPU host logic (world_size 2, gloo): target sharding and the global top-k merge."""
import os
import subprocess
import sys
import textwrap

import numpy as np

from prophylo_b200 import parallel
from prophylo_b200.capi import HIT_DTYPE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fake_hits(rng, nq, nt, t0):
    h = np.zeros(nq * nt, dtype=HIT_DTYPE)
    h["q"] = np.repeat(np.arange(nq), nt)
    h["t"] = np.tile(np.arange(nt), nq)
    h["score"] = rng.choice([0.0, 1e-30, 1e-5, 0.01, 0.5, 1.0], size=h.size) * rng.choice([1.0, 0.5], size=h.size)
    h["depth"] = rng.integers(0, 100, size=h.size)
    return h


def test_shard_range_covers_everything():
    for n, w in [(10, 1), (10, 3), (1000000, 8), (7, 8)]:
        r = [parallel.shard_range(n, i, w) for i in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_merge_topk_equals_global_sort():
    rng = np.random.default_rng(0)
    nq, k = 5, 7
    parts = []
    for r, nt in enumerate([30, 11, 50]):
        h = fake_hits(rng, nq, nt, 0)
        h["t"] += 100 * r
        loc = np.concatenate([np.sort(h[h["q"] == q], order=["score", "t"])[:k] for q in range(nq)])
        parts.append((h, loc))
    merged = parallel.merge_topk([p[1] for p in parts], k, nq)
    allh = np.concatenate([p[0] for p in parts])
    for q in range(nq):
        exp = np.sort(allh[allh["q"] == q], order=["score", "t"])[:k]
        got = merged[merged["q"] == q]
        assert np.array_equal(got, exp)


def test_global_topk_two_ranks_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {ROOT!r})
        import numpy as np, torch.distributed as dist
        from prophylo_b200 import parallel
        from prophylo_b200.capi import HIT_DTYPE
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        nq, k, nt = 4, 5, 40
        rng = np.random.default_rng(123)   # same stream on both ranks: build the global problem, then shard it
        h = np.zeros(nq * nt, dtype=HIT_DTYPE)
        h["q"] = np.repeat(np.arange(nq), nt); h["t"] = np.tile(np.arange(nt), nq)
        h["score"] = rng.choice([0.0, 1e-9, 0.3, 1.0], size=h.size) * rng.random(h.size).round(1)
        t0, t1 = parallel.shard_range(nt, rank, world)
        mine = h[(h["t"] >= t0) & (h["t"] < t1)].copy()
        mine["t"] -= t0
        loc = np.concatenate([np.sort(mine[mine["q"] == q], order=["score", "t"])[:k] for q in range(nq)])
        got = parallel.global_topk(loc, t0, k, nq)
        exp = np.concatenate([np.sort(h[h["q"] == q], order=["score", "t"])[:k] for q in range(nq)])
        assert np.array_equal(got, exp), (rank, got, exp)
        dist.barrier(); dist.destroy_process_group()
        print("ok", rank)
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", str(script)], capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_sliced_global_topk_two_ranks_gloo(tmp_path):
    """The cheap merge: every rank merges only its slice of the queries (all-to-all of the list slices); the slices together
    equal the global sort.  Uneven query slices (5 queries over 2 ranks) and short lists (counts < k) included."""
    script = tmp_path / "w2.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys
        sys.path.insert(0, {ROOT!r})
        import numpy as np, torch.distributed as dist
        from prophylo_b200 import parallel
        from prophylo_b200.capi import HIT_DTYPE
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        nq, k, nt = 5, 6, 9
        rng = np.random.default_rng(77)
        h = np.zeros(nq * nt, dtype=HIT_DTYPE)
        h["q"] = np.repeat(np.arange(nq), nt); h["t"] = np.tile(np.arange(nt), nq)
        h["score"] = rng.choice([0.0, 1e-9, 0.3, 1.0], size=h.size) * rng.random(h.size).round(1)
        t0, t1 = parallel.shard_range(nt, rank, world)
        mine = h[(h["t"] >= t0) & (h["t"] < t1)].copy()
        mine["t"] -= t0
        lists = np.zeros((nq, k), dtype=HIT_DTYPE)
        counts = np.zeros(nq, dtype=np.uint32)
        for q in range(nq):
            loc = np.sort(mine[mine["q"] == q], order=["score", "t"])[:k]     # the shard holds fewer than k targets: short lists
            lists[q, :loc.size] = loc
            counts[q] = loc.size
        q0, got = parallel.global_topk_sliced_host(lists, counts, t0, k, nq)
        a, b = parallel.shard_range(nq, rank, world)
        assert q0 == a
        exp = np.concatenate([np.sort(h[h["q"] == q], order=["score", "t"])[:k] for q in range(a, b)])
        assert np.array_equal(got, exp), (rank, got, exp)
        dist.barrier(); dist.destroy_process_group()
        print("ok", rank)
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29613", str(script)], capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_balanced_query_shards_cover_and_balance():
    """parallel.balanced_query_shards: every query exactly once, rank sizes within one, general presence planes (the queries
    that cost twice as much in the screen) within one per rank, |Y| sums close -- where contiguous slices of the same random
    set differ by several percent."""
    from prophylo_b200 import parallel, synth

    for G, nq, world in ((4096, 1000, 8), (1000, 37, 3), (2048, 5, 8)):
        P, Y, p = synth.make_queries(G, nq, 77 + G, "mixed")
        shards = parallel.balanced_query_shards(P, Y, G, world)
        assert len(shards) == world
        allq = np.concatenate(shards)
        assert np.array_equal(np.sort(allq), np.arange(nq))
        assert all(np.all(np.diff(s) > 0) for s in shards if s.size > 1)
        sizes = [s.size for s in shards]
        assert max(sizes) - min(sizes) <= 1
        general = ~(synth.unpack_bits(P, G) == 1).all(axis=1)
        cnt = [int(general[s].sum()) for s in shards]
        assert max(cnt) - min(cnt) <= 1, cnt
        if nq >= 1000:
            ny = synth.unpack_bits(Y, G).sum(axis=1)
            tot = np.array([ny[s].sum() for s in shards], dtype=np.float64)
            assert tot.max() / tot.min() < 1.05
