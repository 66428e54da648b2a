#!/usr/bin/env python
"""bench.py -- PPP profile-pairs scored per second (BASELINE.json metric) on N B200s of one node.

  python bench.py --gpus 1 --steps K --warmup W                 # this framework (CUDA through the C ABI)
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...                          # the reference's own Perl path on the host cores

Workload = config 3 of BASELINE.json, the configuration the metric is quoted on: 10,000 query profiles x 1,000,000 target
profiles all-pairs at G = 4096 genomes (1e10 pairs), per-query top-k with k = 100.  The target
set is 8 blocks of 125,000 (block b seeded SEED_BASE + 103 + 1000 b), so every N scores the SAME million targets: N = 1
is the whole configuration on one GPU, N = 8 is the configuration as BASELINE names it ("scaling": "strong").  The query x
target tile partitioner cuts the QUERY axis when there are enough queries per GPU (rank r scores queries [r Q / N,
(r + 1) Q / N) against all targets: every per-query cost -- threshold seeding, staircases, list merges -- divides by N and
the ranks' lists are final, no exchange at all) and the TARGET axis otherwise (rank r holds blocks [8 r / N, 8 (r + 1) / N),
global top-k merged over NVLink).  The per-GPU-share form of round 1 (125,000 targets per GPU, target-sharded, work
growing with N, merge included) is reported beside it under "weak_scaling".  A step is one pass of the scan over all pairs.

One JSON line on rank 0; see the task contract for the keys.  `value` is measured with inputs resident in HBM (wall time
of run + result collection bracketed by synchronizes, max over ranks; the library's CUDA-event time is reported beside
it); `e2e` goes through ppp_score() with pinned HOST buffers (targets + queries H2D and hits D2H inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from prophylo_b200 import synth  # noqa: E402

METRIC = "ppp_profile_pairs_per_sec_G4096"
UNIT = "pairs/s"
NBLOCKS = 8


def workload(args):
    return dict(G=args.genomes, nq=args.queries, nt=args.targets, k=args.topk)


def make_queries(w):
    return synth.make_queries(w["G"], w["nq"], synth.SEED_BASE + 3, "mixed")


def target_block(w, b, Y):
    return synth.make_targets(w["G"], w["nt"] // NBLOCKS, synth.SEED_BASE + 103 + 1000 * b, plant_from=Y, plant_frac=0.01)


def rank_blocks(rank, world):
    """blocks of the target set held by `rank`: contiguous, NBLOCKS / world each (world must divide NBLOCKS)"""
    per = NBLOCKS // world
    return list(range(rank * per, (rank + 1) * per))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arms
def _sample_inputs(w, nq, nt):
    P, Y, p = synth.make_queries(w["G"], nq, synth.SEED_BASE + 3, "mixed")
    T = synth.make_targets(w["G"], nt, synth.SEED_BASE + 103, plant_from=Y, plant_frac=0.01)
    return P, Y, p, T


def cpu_port_throughput(w, budget_s, nthreads):
    """The C port of the reference path (oracle/: restatement of ppp.pm:130-245 + Math::Cephes bdtrc, pinned bit-for-bit
    against the reference run under perl) over a bounded sample of the same synthetic workload, all host threads."""
    from oracle import oracle

    P, Y, p, T = _sample_inputs(w, 16, 64)
    t0 = time.perf_counter()
    oracle.score_bits_batch(P[:4], Y[:4], p[:4], T[:32], w["G"], nthreads=nthreads)  # calibration
    rate = 128 / max(time.perf_counter() - t0, 1e-6)
    npairs = int(min(max(rate * budget_s, 256), 16 * 64 * 1024))
    nq = 16
    nt = max(16, npairs // nq)
    T = synth.make_targets(w["G"], nt, synth.SEED_BASE + 103, plant_from=Y, plant_frac=0.01)
    t0 = time.perf_counter()
    _, steps = oracle.score_bits_batch(P, Y, p, T, w["G"], nthreads=nthreads)
    dt = time.perf_counter() - t0
    return dict(value=nq * nt / dt, unit=UNIT, cores=nthreads, kind="port",
                sample=f"{nq} queries x {nt} targets of the same synthetic workload ({nq * nt} pairs, {steps} bdtrc sweep steps, {dt:.1f} s)"), dt


def perl_reference_available():
    return (os.path.exists(os.path.join(ROOT, "baseline", "_ref", "lib", "PhyloProf", "Algorithm", "ppp.pm"))
            and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "Cephes_shim.so")) and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "Cephes.so")))


def perl_reference_throughput(w, budget_s, nproc):
    """The reference ITSELF: the unmodified lib/PhyloProf/Algorithm/ppp.pm (shipped to the box in git-ignored baseline/_ref/,
    BASELINE.md section 3) with Math::Cephes::bdtrc bound to the reference's own binary, one forked perl per host core
    (oracle/time_reference.pl), on a bounded sample of the same synthetic workload.  Every result is cross-checked against
    the C port: the two CPU baselines are the same function."""
    from oracle import oracle

    G = w["G"]
    nq, nt = 16, max(16, 96 * nproc)  # ~6-10 s of wall on the box (16 cores); the workers stop at the budget if it is less
    P, Y, p, T = _sample_inputs(w, nq, nt)
    Pb, Yb, Tb = synth.unpack_bits(P, G), synth.unpack_bits(Y, G), synth.unpack_bits(T, G)
    cases = {"queries": [{str(g + 1): int(Yb[q, g]) for g in range(G) if Pb[q, g]} for q in range(nq)], "probs": [float(x) for x in p],
             "targets": [[str(g + 1) for g in np.nonzero(Tb[t])[0]] for t in range(nt)]}
    with tempfile.TemporaryDirectory() as d:
        json.dump(cases, open(f"{d}/cases.json", "w"))
        env = dict(os.environ, REFERENCE_ROOT=os.path.join(ROOT, "baseline", "_ref"))
        subprocess.run(["perl", os.path.join(ROOT, "oracle", "time_reference.pl"), f"{d}/cases.json", str(nproc), str(budget_s), f"{d}/out.json"],
                       env=env, check=True, timeout=budget_s * 6 + 120)
        out = json.load(open(f"{d}/out.json"))
    ref, _ = oracle.score_bits_batch(P, Y, p, T, G, nthreads=nproc)
    bad = 0
    for key, v in out["results"].items():
        q, t = (int(x) for x in key.split(","))
        r = ref[q, t]
        bad += (float(v[0]), v[1], v[2], v[3]) != (float(r["score"]), int(r["yes"]), int(r["number"]), int(r["depth"]))
    return dict(value=out["pairs"] / out["seconds"], unit=UNIT, cores=nproc, kind="reference",
                sample=f"{out['pairs']} (query, target) pairs of the same synthetic workload ({nq} queries x {nt} targets drawn; "
                       f"{out['seconds']:.1f} s wall; unmodified ppp.pm under perl, {nproc} forked workers)",
                mismatches_vs_port=int(bad)), out["seconds"]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    w = workload(args)
    ncpu = os.cpu_count() or 1
    use_perl = perl_reference_available()
    vals = []
    info = None
    budget = max(2.0, min(15.0, 100.0 / max(1, args.steps + args.warmup)))
    for i in range(args.warmup + args.steps):
        info, dt = perl_reference_throughput(w, budget, ncpu) if use_perl else cpu_port_throughput(w, budget, ncpu)
        if i >= args.warmup:
            vals.append((info["value"], dt))
    v = float(np.mean([x[0] for x in vals]))
    info["value"] = v
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.mean([x[1] for x in vals])), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"config3 sample: {w['nq']} queries x {w['nt']} targets x G={w['G']} all-pairs; the reference's CPU path timed on a "
                               f"bounded sample of the same synthetic pairs (see cpu_baseline.sample)", "genomes": w["G"]},
        "cpu_baseline": info,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if use_perl:  # the C port beside it (the stronger CPU baseline: about 19 x the Perl rate per core)
        line["cpu_baseline_port"] = cpu_port_throughput(w, 8.0, ncpu)[0]
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------ other configurations
def _timed(ctx, flt, reps=3):
    best = None
    for _ in range(reps):
        ctx.run(flt)
        s = ctx.stats()
        if best is None or s["total_device_ms"] < best["total_device_ms"]:
            best = s
    return best


def other_configs(ctx, torch):
    """BASELINE configs 2, 4, 5 and the dense (every pair reported) form at G = 4096 on one GPU: device time with inputs
    resident (library CUDA events, best of 3), plus config 2 end to end through host buffers.  Extra information beside the
    headline workload; not part of `value`."""
    from prophylo_b200 import capi

    out = {}
    # config 2: the single-query ppp.pl shape, 1 query x 1,000,000 targets x 1024 genomes
    G, nt = 1024, 1_000_000
    P, Y, p = synth.make_queries(G, 1, synth.SEED_BASE + 2, "single")
    T = synth.make_targets(G, nt, synth.SEED_BASE + 102, plant_from=Y, plant_frac=0.01)
    Tt = torch.from_numpy(T).pin_memory()
    buf = torch.empty(nt * capi.HIT_DTYPE.itemsize, dtype=torch.uint8).pin_memory().numpy().view(capi.HIT_DTYPE)
    ctx.set_targets_bits(Tt.numpy(), G)
    ctx.set_queries(P, Y, p)
    s = _timed(ctx, None, 5)
    used = ctx.counter("used_table")
    e2e = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.set_targets_bits(Tt.numpy(), G, streamed=True)  # the sweep of a piece overlaps the upload of the next one
        ctx.score(P, Y, p, None, out=buf)
        e2e.append(time.perf_counter() - t0)
    alg_bytes = nt * (G // 8 + 24)  # one target plane + a 24-byte result per pair
    out["config2_dense_1x1M_G1024"] = {
        "pairs_per_s_resident": nt / (s["total_device_ms"] * 1e-3), "ms": s["total_device_ms"],
        "pairs_per_s_e2e_host_buffers": nt / float(np.median(e2e)), "mode": "exact table" if used else "general sweep",
        "hbm_GBps_algorithmic": alg_bytes / (s["total_device_ms"] * 1e-3) / 1e9, "hits_out_bytes": int(buf.nbytes), "targets_in_bytes": int(T.nbytes)}
    s = _timed(ctx, capi.make_filter(capi.FILTER_TOPK, k=100))
    out["config2_top100"] = {"ms": s["total_device_ms"], "pairs_per_s_resident": nt / (s["total_device_ms"] * 1e-3)}
    # config 4: ppp2d shape, 2,000 nested prefix queries x 500,000 targets x 2048 genomes, best 2 per query
    G, nq, nt = 2048, 2000, 500_000
    P, Y, p = synth.make_queries(G, nq, synth.SEED_BASE + 4, "prefix")
    T = synth.make_targets(G, nt, synth.SEED_BASE + 104, plant_from=Y, plant_frac=0.01)
    ctx.set_targets_bits(T, G)
    ctx.set_queries(P, Y, p)
    s = _timed(ctx, capi.make_filter(capi.FILTER_TOPK, k=2))
    out["config4_ppp2d_top2"] = {"shape": f"{nq} x {nt} x G={G}", "ms": s["total_device_ms"], "pairs_per_s_resident": nq * nt / (s["total_device_ms"] * 1e-3),
                                 "prefilter_ms": ctx.counter("us_prefilter") / 1e3, "survivors": ctx.counter("survivors")}
    # config 5: TIGRFAM-scale all-pairs, 4,500 x 4,500 x 8192 genomes: ppp_cutoff threshold filter and top-100
    G, nq = 8192, 4500
    P, Y, p = synth.make_queries(G, nq, synth.SEED_BASE + 5, "family")
    ctx.set_targets_bits(Y, G)
    ctx.set_queries(P, Y, p)
    s = _timed(ctx, capi.make_filter(capi.FILTER_THRESH_LOG10, thresh=-0.0005), 2)
    out["config5_thresh"] = {"shape": f"{nq} x {nq} x G={G}", "ms": s["total_device_ms"], "pairs_per_s_resident": nq * nq / (s["total_device_ms"] * 1e-3),
                             "hits": ctx.result_count(), "rank_ms": ctx.counter("us_merge") / 1e3}
    s = _timed(ctx, capi.make_filter(capi.FILTER_TOPK, k=100), 2)
    out["config5_top100"] = {"ms": s["total_device_ms"], "pairs_per_s_resident": nq * nq / (s["total_device_ms"] * 1e-3),
                             "prefilter_ms": ctx.counter("us_prefilter") / 1e3, "survivors": ctx.counter("survivors")}
    # dense at G = 4096: every pair's (score, yes, number, depth) produced (no filter), config 3's query mix
    G, nq, nt = 4096, 1000, 20_000
    P, Y, p = synth.make_queries(G, nq, synth.SEED_BASE + 3, "mixed")
    T = synth.make_targets(G, nt, synth.SEED_BASE + 103, plant_from=Y, plant_frac=0.01)
    ctx.set_targets_bits(T, G)
    ctx.set_queries(P, Y, p)
    ctx.set_option("table_mode", 0)
    try:
        s = _timed(ctx, None, 2)
    finally:
        ctx.set_option("table_mode", -1)
    out["dense_G4096"] = {"shape": f"{nq} x {nt} x G={G}, every pair reported", "ms": s["total_device_ms"],
                          "pairs_per_s_resident": nq * nt / (s["total_device_ms"] * 1e-3), "candidates_per_pair": s["candidates"] / (nq * nt),
                          "exact_pairs": s["exact_pairs"]}
    return out


# ------------------------------------------------------------------------------------------------ this framework
def run_ours(args):
    import torch
    import torch.distributed as dist

    from prophylo_b200 import capi, parallel

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if NBLOCKS % world:
        raise SystemExit(f"bench.py: the target set is {NBLOCKS} blocks; --gpus must divide it")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    w = workload(args)
    G, nq, nt_total, k = w["G"], w["nq"], w["nt"], w["k"]
    P, Y, p = make_queries(w)
    partition = args.partition if args.partition != "auto" else ("queries" if world > 1 and nq // world >= 512 else "targets")
    block_nt = nt_total // NBLOCKS
    if partition == "queries":   # rank r: its share of the queries (balanced by presence-plane class and |Y|) against ALL targets
        blocks = list(range(NBLOCKS))
        qidx = parallel.balanced_query_shards(P, Y, G, world)[rank]
    else:                        # rank r: all queries against its blocks of the targets
        blocks = rank_blocks(rank, world)
        qidx = np.arange(nq)
    q0, q1 = 0, len(qidx)
    T = np.concatenate([target_block(w, b, Y) for b in blocks])
    nt = T.shape[0]

    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()

    hold = [pinned(x) for x in (P, Y, p, T, P[qidx], Y[qidx], p[qidx])]
    Pp, Yp, pp, Tp, Pq, Yq, pq = [h[1] for h in hold]   # ..q: this rank's queries of the headline run
    ctx = capi.Context(local_rank)
    ctx.set_option("sweep_warps", args.sweep_warps)
    flt = capi.make_filter(capi.FILTER_TOPK, k=k)
    out_t = torch.empty(nq * min(k, nt) * capi.HIT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    out_np = out_t.numpy().view(capi.HIT_DTYPE)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    merger = parallel.SlicedTopkMerge(ctx, nq, k, world, rank, dev) if world > 1 else None

    def collect(offset_of_rank):
        """results of the last run on the host.  Lists that are final on this rank (one GPU; query-partitioned run): the device
        top-k lists are copied into the pinned buffer.  Target-partitioned run on N GPUs: the path's only exchange step, the
        global top-k merge -- rank r receives every rank's device-resident lists of ITS slice of the queries (NCCL all-to-all
        over NVLink), merges them on the device and copies its slice of the global lists to its host buffer (the result of an
        N-process job is the N slices)."""
        if offset_of_rank is None:
            return ctx.fetch(out_np)
        return merger.run(offset_of_rank)

    def timed_steps(nsteps, offset_of_rank, sample_clocks):
        sampler = ClockSampler(local_rank) if sample_clocks else None
        barrier()
        if sampler:
            sampler.start()
        r = dict(dev_ms=[], sweep_ms=[], launches=0, wall=0.0, pf_us=[], pf_pairs=[], survs=[])
        for _ in range(nsteps):
            flush.zero_()  # flush L2 between timed iterations (untimed)
            barrier()
            t0 = time.perf_counter()
            ctx.run(flt)
            collect(offset_of_rank)
            torch.cuda.synchronize()
            r["wall"] += time.perf_counter() - t0
            st = ctx.stats()
            r["dev_ms"].append(st["total_device_ms"])
            r["sweep_ms"].append(st["sweep_kernel_ms"])
            r["launches"] += st["kernel_launches"] + (merger.launches_per_run if offset_of_rank is not None else 0)
            r["pf_us"].append(ctx.counter("us_prefilter"))
            r["pf_pairs"].append(ctx.counter("prefilter_pairs"))
            r["survs"].append(ctx.counter("survivors"))
        barrier()
        r["clocks"] = sampler.stop() if sampler else None
        t = torch.tensor([1e3 * r["wall"] / nsteps, float(np.mean(r["dev_ms"])), float(np.mean(r["sweep_ms"]))], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        r["step_ms"], r["dev_ms_mean"], r["sweep_ms_mean"] = [float(x) for x in t.cpu()]
        return r

    # ---- the headline: the whole configuration over the N GPUs (resident inputs)
    ctx.set_targets_bits(Tp, G)
    ctx.set_queries(Pq, Yq, pq)
    strong_off = None if (world == 1 or partition == "queries") else (lambda r: rank_blocks(r, world)[0] * block_nt)
    for _ in range(args.warmup):
        ctx.run(flt)
        collect(strong_off)
    if args.profile_step:
        flush.zero_()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        ctx.run(flt)
        collect(strong_off)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    main = timed_steps(args.steps, strong_off, True)
    clocks = main["clocks"]
    value = nq * nt_total / (main["step_ms"] * 1e-3)
    full_pairs = ctx.counter("full_pairs")
    phase = {n: ctx.counter("us_" + n) for n in ("sweep", "prefilter", "exact", "staircase", "merge")}
    spec = {"speculative_rank": ctx.counter("spec_rank"), "queries_rescanned": ctx.counter("spec_failed")}

    # ---- end to end through the one-call C ABI entry with host buffers (targets + queries H2D, hits D2H)
    e2e_wall = 0.0
    e2e_score_ms, e2e_dev_ms = [], []
    nsteps_e2e = max(1, min(args.steps, 3))
    # One GPU, or target-partitioned: the rank's planes are uploaded with the streamed entry point (the scan of the first pieces
    # overlaps the copy of the later ones).  Query-partitioned on N GPUs: every GPU needs ALL targets, and N full uploads share the
    # host's memory and PCIe bandwidth (8 GPUs: 55 ms for an upload that takes 10 ms alone) -- each rank uploads 1/N of the planes,
    # the ranks exchange their slices over NVLink (NCCL all-gather) and hand the gathered device buffer to the library
    # (ppp_set_targets_bits_device).  Inputs still start in host buffers; all of it is inside the timed region.  At N = 2 the
    # streamed full upload still hides behind the scan (measured: 167.4 ms against 171.4 ms with the exchange), so N >= 4 only.
    gather_targets = world >= 4 and partition == "queries" and nt % world == 0
    if gather_targets:
        sl = nt // world
        Tslice_host = torch.from_numpy(Tp[rank * sl:(rank + 1) * sl].view(np.int32))  # view of the pinned buffer (NCCL has no uint32)
        Tslice_dev = torch.empty(Tslice_host.shape, dtype=Tslice_host.dtype, device=dev)
        Tfull_dev = torch.empty((nt, Tp.shape[1]), dtype=Tslice_host.dtype, device=dev)

    def upload_targets():
        if gather_targets:
            Tslice_dev.copy_(Tslice_host, non_blocking=True)
            dist.all_gather_into_tensor(Tfull_dev, Tslice_dev)
            torch.cuda.current_stream().synchronize()
            ctx.set_targets_bits_device(Tfull_dev.data_ptr(), nt, G)
        else:
            ctx.set_targets_bits(Tp, G, streamed=True)

    upload_targets()
    ctx.score(Pq, Yq, pq, flt, out=out_np)
    for _ in range(nsteps_e2e):
        flush.zero_()
        barrier()
        t0 = time.perf_counter()
        upload_targets()
        hits = ctx.score(Pq, Yq, pq, flt, out=out_np)
        t1 = time.perf_counter()
        if strong_off is not None:
            collect(strong_off)
        torch.cuda.synchronize()
        e2e_wall += time.perf_counter() - t0
        e2e_score_ms.append(1e3 * (t1 - t0))
        e2e_dev_ms.append(ctx.stats()["total_device_ms"])
    barrier()
    e2e_ms = torch.tensor([1e3 * e2e_wall / nsteps_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = nq * nt_total / (float(e2e_ms.item()) * 1e-3)
    h2d = int((Tp.nbytes // world if gather_targets else Tp.nbytes) + Pq.nbytes + Yq.nbytes + pq.nbytes)
    d2h = int(hits.nbytes) if strong_off is None else int(merger.d2h_bytes)

    # ---- the per-GPU-share form (round 1's line): 125,000 targets per GPU whatever N is, target-sharded, global merge included;
    # N > 1: afterwards the merged global lists of 64 queries are checked against ONE GPU scoring the concatenated shards
    wb = blocks.index(rank) if partition == "queries" and world > 1 else 0  # a different block on every rank
    Tw = Tp[wb * block_nt:(wb + 1) * block_nt]
    ctx.set_targets_bits(Tw, G)
    ctx.set_queries(Pp, Yp, pp)
    weak_off = None if world == 1 else (lambda r: r * block_nt)
    for _ in range(2):
        ctx.run(flt)
        collect(weak_off)
    wk = timed_steps(max(2, min(args.steps, 5)), weak_off, False)
    weak = {"scaling": "weak", "partition": "targets", "targets_per_gpu": block_nt, "value": world * nq * block_nt / (wk["step_ms"] * 1e-3),
            "ms_per_step": wk["step_ms"], "device_ms_per_step": wk["dev_ms_mean"]}
    selfcheck = None
    if world > 1:
        selfcheck = parallel.check_global_lists(ctx, merger, weak_off, Pp, Yp, pp, Tw, G, k, rank, world, dev, n_check=64)
        ctx.set_targets_bits(Tp, G)
        ctx.set_queries(Pq, Yq, pq)

    if rank == 0:
        pf_us_mean, pf_pairs_mean, surv_mean = float(np.mean(main["pf_us"])), float(np.mean(main["pf_pairs"])), float(np.mean(main["survs"]))
        pf_ms = pf_us_mean * 1e-3
        words = ((G + 1023) // 1024) * 32
        Pbits, Ybits = synth.unpack_bits(Pq, G), synth.unpack_bits(Yq, G)
        is_all = Pbits.all(axis=1)
        is_pyy = ~is_all & (Pbits == Ybits).all(axis=1)
        frac_all, frac_pyy = float(np.mean(is_all)), float(np.mean(is_pyy))
        # POPC the dominant kernel (ppp_prefilter2_kernel, ppp_prefilter.cu) EXECUTES per pair: carry-save adders fold a
        # 4-word chunk into 3 POPC (8 words -> 4 at G = 8192); the first chunk is screened word by word; all-ones queries take
        # `number` from a per-target depth prefix computed once per 32-query tile; queries with P == Y need no `number` at all
        # (it equals yes); general queries count T & P as well
        ch = 8 if words == 256 else 4
        nch, per_chunk = words // ch, (4 if ch == 8 else 3)
        popc_all = per_chunk * (nch - 1) + ch + ch / 8.0 + per_chunk * nch / 32.0
        popc_pyy = per_chunk * nch
        popc_gen = 2 * per_chunk * (nch - 1) + 2 * ch
        popc_exec = frac_all * popc_all + frac_pyy * popc_pyy + (1.0 - frac_all - frac_pyy) * popc_gen
        # SURVEY.md section 8(d) counts the straightforward algorithm: 2W POPC per pair when P is all ones, 3W otherwise
        popc_survey = frac_all * 2 * words + (1.0 - frac_all) * 3 * words
        nsm = ctx.counter("num_sms")
        peaks = {name: ctx.pipe_peak(name) for name in ("popc", "lop3", "iadd", "dfma")}  # measured now, under this run's clocks
        clk = (clocks.get("sm_mhz") or 1965.0) * 1e6
        popc_peak_analytic = nsm * 16 * clk
        pairs_rate = pf_pairs_mean / (pf_ms * 1e-3) if pf_ms > 0 else 0.0
        popc_rate = pairs_rate * popc_exec
        # DRAM traffic of the dominant kernel per step: from the committed ncu capture of exactly this command, if there is one
        traffic, traffic_step, traffic_src = None, None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")))
            if tj.get("workload") == [G, nq, nt_total, k, world]:
                traffic_step, traffic_src = tj["dram_bytes_per_step"], tj["source"]
                traffic = traffic_step / tj["launches_per_step"]  # per launch, like the contract's `achieved`
        except (OSError, ValueError, KeyError):
            pass
        W_bytes = words * 4
        alg_bytes = nt * W_bytes + (q1 - q0) * 2 * W_bytes + 4 * surv_mean
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": main["step_ms"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"config3 (BASELINE: 10k x 1M all-pairs at 4096 genomes): {nq} queries x {nt_total} targets x G={G}, per-query "
                                   f"top-k k={k}; {world} GPU(s), partition = {partition}: {q1 - q0} queries x {nt} targets each", "genomes": G, "queries": nq,
                       "targets": nt_total, "partition": partition + (" (dealt over the GPUs by presence-plane class and |Y|: parallel.balanced_query_shards)" if partition == "queries" else ""), "queries_per_gpu": q1 - q0, "targets_per_gpu": nt, "filter": f"topk{k}", "l2": "flushed between timed iterations (256 MiB write)",
                       "timing": "wall of run + result collection bracketed by synchronize, max over ranks; CUDA-event time of the library's "
                                 "stream in device_ms_per_step"},
            "device_ms_per_step": main["dev_ms_mean"], "sweep_kernel_ms_per_step": main["sweep_ms_mean"],
            "phase_us_last_step": phase, "full_pairs_last_step": full_pairs,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "rank0_ms": {"upload_and_score_call": float(np.mean(e2e_score_ms)), "device_part_of_it": float(np.mean(e2e_dev_ms))},
                    "note": ("every rank uploads 1/N of the target planes, NCCL all-gather over NVLink, ppp_set_targets_bits_device; h2d_bytes_per_step is per rank"
                             if gather_targets else "targets uploaded with ppp_set_targets_bits_streamed: the copy overlaps the scan")},
            "gpu_launches": main["launches"],
            "roofline": {"bound": "int_pipe", "pipe": "xu_popc", "kernel": "ppp_prefilter2_kernel", "achieved": popc_rate / 1e12,
                         "peak": peaks["popc"] / 1e12, "unit": "TPOPC/s", "frac": popc_rate / peaks["popc"] if peaks["popc"] else None,
                         "traffic": traffic, "traffic_per_step": traffic_step, "traffic_source": traffic_src, "algorithmic_hbm_bytes_per_step": alg_bytes,
                         "hbm_GBps_algorithmic": alg_bytes / (pf_ms * 1e-3) / 1e9 if pf_ms > 0 else None,
                         "kernel_ms_per_step": pf_ms, "kernel_share_of_step": pf_ms / main["step_ms"],
                         "popc_executed_per_pair": popc_exec, "popc_per_pair_survey_8d": popc_survey,
                         "frac_vs_survey_count": pairs_rate * popc_survey / peaks["popc"] if peaks["popc"] else None,
                         "pairs_per_s_in_kernel": pairs_rate, "all_ones_presence_fraction": frac_all, "p_equals_y_fraction": frac_pyy,
                         "peak_source": "ppp_pipe_peak microbenchmark run in this process after the timed steps (popc.b32 chains, 64 warps/SM)",
                         "peak_analytic": popc_peak_analytic / 1e12,
                         "peak_analytic_source": f"{nsm} SMs x 16 POPC/clk/SM x {clk / 1e6:.0f} MHz (sampled SM clock)",
                         "note": "both operands are reused on chip (targets in L2, query tiles in shared memory), so the kernel is bound by the "
                                 "POPC (XU) pipe, not by HBM; frac = executed POPC / measured pipe peak; frac_vs_survey_count uses the "
                                 "straightforward algorithm's POPC count and can exceed 1 because carry-save adders and the shared depth "
                                 "prefix remove POPCs"},
            "pipe_peaks_measured": {name: v / 1e12 for name, v in peaks.items()}, "pipe_peaks_unit": "T lane-ops/s",
            "topk_schedule": dict(spec, survivors_fully_evaluated=int(surv_mean)),
            "weak_scaling": weak,
        }
        if selfcheck is not None:
            line["multi_gpu_selfcheck"] = selfcheck
        if world == 1:
            ncpu = os.cpu_count() or 1
            port, _ = cpu_port_throughput(w, 12.0, ncpu)
            if perl_reference_available():
                line["cpu_baseline"], _ = perl_reference_throughput(w, 12.0, ncpu)
                line["cpu_baseline_port"] = port
            else:
                line["cpu_baseline"] = port
            try:
                line["other_configs"] = other_configs(ctx, torch)
            except Exception as e:  # the headline line must not depend on the extra configurations
                line["other_configs"] = {"error": f"{type(e).__name__}: {e}"}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genomes", type=int, default=4096)
    ap.add_argument("--queries", type=int, default=10000)
    ap.add_argument("--targets", type=int, default=1000000, help="targets of the whole job (8 blocks)")
    ap.add_argument("--topk", type=int, default=100)
    ap.add_argument("--sweep-warps", type=int, default=16)
    ap.add_argument("--profile-step", action="store_true",
                    help="for ncu --profile-from-start off: after the warm-up, run ONE step between cudaProfilerStart/Stop and exit "
                         "(no timing, no JSON line)")
    ap.add_argument("--partition", default="auto", choices=["auto", "queries", "targets"],
                    help="axis the N GPUs split: queries (lists final per rank, no exchange) or targets (global top-k merge); auto = queries when every GPU gets >= 512")
    args = ap.parse_args()
    if args.targets % NBLOCKS:
        raise SystemExit(f"--targets must be a multiple of {NBLOCKS}")
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
