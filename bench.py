#!/usr/bin/env python
"""bench.py -- PPP profile-pairs scored per second (BASELINE.json metric) on N B200s of one node.

  python bench.py --gpus 1 --steps K --warmup W                 # this framework (CUDA through the C ABI)
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...                          # the reference's CPU path (oracle port, all host threads)

Workload (config 3 of BASELINE.json, the configuration the metric is quoted on: 4096 genomes): 10,000 query
profiles x 1,000,000 target profiles all-pairs at G = 4096, sharded over the target axis -- 125,000 targets per GPU,
so N = 8 is the full configuration and N = 1 its per-GPU share ("scaling": "weak").  Filter: per-query top-k, k = 100
(SURVEY.md section 8d config 3).  A step is one pass of the scan over all resident pairs.

One JSON line on rank 0; see the task contract for the keys.  `value` is measured with inputs resident in HBM and
timed with the library's CUDA events on its own stream (max over ranks); `e2e` goes through ppp_score() with pinned
HOST buffers (targets + queries H2D and hits D2H inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from prophylo_b200 import synth  # noqa: E402

METRIC = "ppp_profile_pairs_per_sec_G4096"
UNIT = "pairs/s"


def workload(args):
    return dict(G=args.genomes, nq=args.queries, nt_per_gpu=args.targets_per_gpu, k=args.topk)


def make_inputs(w, rank):
    P, Y, p = synth.make_queries(w["G"], w["nq"], synth.SEED_BASE + 3, "mixed")
    T = synth.make_targets(w["G"], w["nt_per_gpu"], synth.SEED_BASE + 103 + 1000 * rank, plant_from=Y, plant_frac=0.01)
    return P, Y, p, T


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


def cpu_reference_throughput(w, budget_s, nthreads):
    """The reference's CPU path for this workload: oracle port (C restatement of ppp.pm:130-245 + Math::Cephes bdtrc,
    pinned bit-for-bit against the reference run under perl) over a bounded sample of the same pairs, all host threads."""
    from oracle import oracle

    P, Y, p = synth.make_queries(w["G"], 16, synth.SEED_BASE + 3, "mixed")
    T = synth.make_targets(w["G"], 64, synth.SEED_BASE + 103, plant_from=Y, plant_frac=0.01)
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    w = workload(args)
    nthreads = os.cpu_count() or 1
    vals = []
    info = None
    budget = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    for i in range(args.warmup + args.steps):
        info, dt = cpu_reference_throughput(w, budget, nthreads)
        if i >= args.warmup:
            vals.append((info["value"], dt))
    v = float(np.mean([x[0] for x in vals]))
    info["value"] = v
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.mean([x[1] for x in vals])), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"config3 sample: {w['nq']} queries x {w['nt_per_gpu']} targets/GPU x G={w['G']} all-pairs, "
                               f"CPU port timed on a bounded sample (see cpu_baseline.sample)", "genomes": w["G"]},
        "cpu_baseline": info,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def config2_dense(ctx, torch, args):
    """BASELINE config 2 (the single-query ppp.pl shape): 1 query x 1,000,000 targets x 1024 genomes, EVERY pair reported
    (dense).  Few queries x many targets -> exact table mode: scores bit-identical to the reference.  Extra information
    beside the headline workload; not part of `value`."""
    from prophylo_b200 import capi

    G, nt = 1024, 1000000
    P, Y, p = synth.make_queries(G, 1, synth.SEED_BASE + 2, "single")
    T = synth.make_targets(G, nt, synth.SEED_BASE + 102, plant_from=Y, plant_frac=0.01)
    Tt = torch.from_numpy(T).pin_memory()
    out = torch.empty(nt * capi.HIT_DTYPE.itemsize, dtype=torch.uint8).pin_memory().numpy().view(capi.HIT_DTYPE)
    ctx.set_targets_bits(Tt.numpy(), G)
    ctx.set_queries(P, Y, p)
    for _ in range(2):
        ctx.run(None)
    dev = []
    for _ in range(5):
        ctx.run(None)
        dev.append(ctx.stats()["total_device_ms"])
    used = ctx.counter("used_table")
    e2e = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.set_targets_bits(Tt.numpy(), G)
        ctx.score(P, Y, p, None, out=out)
        e2e.append(time.perf_counter() - t0)
    return {"pairs_per_s_resident": nt / (float(np.median(dev)) * 1e-3), "pairs_per_s_e2e_host_buffers": nt / float(np.median(e2e)),
            "mode": "exact table" if used else "general sweep", "hits_out_bytes": int(out.nbytes), "targets_in_bytes": int(T.nbytes)}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from prophylo_b200 import capi, parallel

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    w = workload(args)
    G, nq, nt, k = w["G"], w["nq"], w["nt_per_gpu"], w["k"]
    P, Y, p, T = make_inputs(w, rank)

    def pinned(a):
        t = torch.from_numpy(a).pin_memory()
        return t, t.numpy()

    hold = [pinned(x) for x in (P, Y, p, T)]
    Pp, Yp, pp, Tp = [h[1] for h in hold]
    ctx = capi.Context(local_rank)
    ctx.set_option("sweep_warps", args.sweep_warps)
    flt = capi.make_filter(capi.FILTER_TOPK, k=k)
    ctx.set_targets_bits(Tp, G)
    ctx.set_queries(Pp, Yp, pp)
    out_t = torch.empty(nq * min(k, nt) * capi.HIT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    out_np = out_t.numpy().view(capi.HIT_DTYPE)
    cnt_t = torch.empty(nq, dtype=torch.int32).pin_memory()
    cnt_np = cnt_t.numpy().view(np.uint32)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def collect():
        """results of the last run on the host.  One GPU: the device top-k lists are copied into the pinned buffer.
        N GPUs: the path's only exchange step, the global top-k merge -- NCCL all-gather of the device-resident per-rank
        lists + one device merge, result on every rank (the per-rank lists are not fetched separately)."""
        if world == 1:
            return ctx.fetch(out_np)
        out, cnt = parallel.global_topk_device(ctx, lambda r: r * nt, k, nq, torch.device("cuda", local_rank), out=out_np, cnt=cnt_np)
        return out

    # ---- resident-input steps (value)
    for _ in range(args.warmup):
        ctx.run(flt)
        collect()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    dev_ms, sweep_ms, launches, wall = [], [], 0, 0.0
    pf_us, pf_pairs, survs = [], [], []
    st = None
    for _ in range(args.steps):
        flush.zero_()  # flush L2 between timed iterations (resident inputs are smaller than L2), untimed
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.run(flt)
        res = collect()
        torch.cuda.synchronize()
        wall += time.perf_counter() - t0
        st = ctx.stats()
        dev_ms.append(st["total_device_ms"])
        sweep_ms.append(st["sweep_kernel_ms"])
        launches += st["kernel_launches"]
        pf_us.append(ctx.counter("us_prefilter"))
        pf_pairs.append(ctx.counter("prefilter_pairs"))
        survs.append(ctx.counter("survivors"))
    barrier()
    clocks = sampler.stop()
    full_pairs = ctx.counter("full_pairs")
    phase = {n: ctx.counter("us_" + n) for n in ("sweep", "prefilter", "exact", "staircase", "merge")}
    pf_us_mean, pf_pairs_mean, surv_mean = float(np.mean(pf_us)), float(np.mean(pf_pairs)), float(np.mean(survs))
    # the step time is the wall time of run+fetch(+merge) bracketed by synchronizes: it contains the device-event time
    step_ms = 1e3 * wall / args.steps
    t = torch.tensor([step_ms, float(np.mean(dev_ms)), float(np.mean(sweep_ms))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms, dev_ms_mean, sweep_ms_mean = [float(x) for x in t.cpu()]
    pairs_per_gpu = nq * nt
    value = world * pairs_per_gpu / (step_ms * 1e-3)

    # ---- end-to-end steps through the one-call C ABI entry with host buffers (targets + queries H2D, hits D2H)
    e2e_wall = 0.0
    nsteps_e2e = max(1, min(args.steps, 3))
    ctx.set_targets_bits(Tp, G)
    ctx.score(Pp, Yp, pp, flt, out=out_np)
    barrier()
    for _ in range(nsteps_e2e):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.set_targets_bits(Tp, G)
        hits = ctx.score(Pp, Yp, pp, flt, out=out_np)
        if world > 1:
            collect()
        torch.cuda.synchronize()
        e2e_wall += time.perf_counter() - t0
    barrier()
    e2e_ms = torch.tensor([1e3 * e2e_wall / nsteps_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * pairs_per_gpu / (float(e2e_ms.item()) * 1e-3)
    h2d = int(Tp.nbytes + Pp.nbytes + Yp.nbytes + pp.nbytes)
    d2h = int(hits.nbytes)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        # Dominant kernel = the thread-per-pair integer pre-filter (ppp_prefilter2_kernel, ppp_prefilter.cu; two launches
        # per target chunk: all-ones presence planes / general).  Its ALGORITHMIC HBM bytes per step: target planes once,
        # query planes once, survivor ids out; both operands are reused on chip, so the kernel is
        # bound by the integer pipes (POPC on the XU pipe: 16 lanes/clk/SM), not by HBM -- reported as pipe_roofline.
        W_bytes = ((G + 1023) // 1024) * 128
        pf_ms = pf_us_mean * 1e-3
        alg_bytes = nt * W_bytes + nq * 2 * W_bytes + 4 * surv_mean
        achieved = alg_bytes / (pf_ms * 1e-3) / 1e9 if pf_ms > 0 else 0.0
        # DRAM traffic of the step's pre-filter launches, from the committed ncu capture of exactly this workload (the
        # default arguments); other shapes have no capture -> null
        default_shape = (G, nq, nt, k) == (4096, 10000, 125000, 100)
        traffic = 266082816 if default_shape else None
        words = W_bytes // 4
        frac_all = float(np.mean(synth.unpack_bits(P, G).all(axis=1)))
        # POPC the kernel EXECUTES per pair: carry-save adders fold a 4-word chunk into 3 POPC (8 words -> 4 at G=8192);
        # the first chunk is screened word by word; all-ones queries take number from a per-target depth prefix that is
        # computed once per 32-query tile; general queries count T&P as well.
        ch = 8 if words == 256 else 4
        nch, per_chunk = words // ch, (4 if ch == 8 else 3)
        popc_all = per_chunk * (nch - 1) + ch + ch / 8.0 + per_chunk * nch / 32.0
        popc_gen = 2 * per_chunk * (nch - 1) + 2 * ch
        popc_exec = frac_all * popc_all + (1.0 - frac_all) * popc_gen
        # SURVEY.md section 8(d) counts the straightforward algorithm: 2W POPC per pair when P is all ones, 3W otherwise
        popc_survey = frac_all * 2 * words + (1.0 - frac_all) * 3 * words
        clk = (clocks.get("sm_mhz") or 1965.0) * 1e6
        nsm = ctx.counter("num_sms")
        popc_peak = nsm * 16 * clk  # POPC issues on the XU pipe: 16 lanes / clk / SM (ncu: sm__inst_executed_pipe_xu)
        pairs_rate = pf_pairs_mean / (pf_ms * 1e-3) if pf_ms > 0 else 0.0
        popc_rate = pairs_rate * popc_exec
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"config3 (10k x 1M all-pairs at 4096 genomes over 8 GPUs): {nq} queries x {nt} targets per GPU x "
                                   f"G={G}, per-query top-k k={k}; N GPUs = {world * nt} targets", "genomes": G, "queries": nq,
                       "targets_per_gpu": nt, "filter": f"topk{k}", "l2": "flushed between timed iterations (256 MiB write)",
                       "timing": "wall of run+fetch bracketed by synchronize; device-event time in device_ms_per_step"},
            "device_ms_per_step": dev_ms_mean, "sweep_kernel_ms_per_step": sweep_ms_mean,
            "phase_us_last_step": phase, "full_pairs_last_step": full_pairs,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "ppp_prefilter2_kernel", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "algorithmic_bytes": alg_bytes,
                         "traffic_source": "profiles/r01_ncu_bench_prefilter2_traffic.txt: dram__bytes_read.sum + dram__bytes_write.sum "
                                           "summed over the 14 ppp_prefilter2 launches of one step of this workload (ncu, same command)"
                                           if traffic else None,
                         "kernel_ms_per_step": pf_ms, "kernel_share_of_step": pf_ms / step_ms,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                         "note": "operands are reused on chip (targets resident in L2, query tile in shared memory): HBM is not the "
                                 "bound of this kernel, the POPC pipe is -- see pipe_roofline"},
            "pipe_roofline": {"bound": "xu_popc", "kernel": "ppp_prefilter2_kernel", "achieved": popc_rate / 1e12, "peak": popc_peak / 1e12,
                              "unit": "TPOPC/s", "frac": popc_rate / popc_peak, "popc_executed_per_pair": popc_exec,
                              "popc_per_pair_survey_8d": popc_survey, "frac_vs_survey_count": pairs_rate * popc_survey / popc_peak,
                              "pairs_per_s_in_kernel": pairs_rate, "all_ones_presence_fraction": frac_all,
                              "peak_source": f"{nsm} SMs x 16 POPC/clk/SM x {clk / 1e6:.0f} MHz (sampled SM clock)",
                              "note": "frac = executed POPC / pipe peak (a utilisation); frac_vs_survey_count uses the straightforward "
                                      "algorithm's POPC count and can exceed 1 because carry-save adders and the shared depth prefix "
                                      "remove POPCs"},
            "topk_schedule": {"speculative_rank": ctx.counter("spec_rank"), "queries_rescanned": ctx.counter("spec_failed"),
                              "survivors_fully_evaluated": int(surv_mean)},
        }
        if world == 1:
            info, _ = cpu_reference_throughput(w, 15.0, os.cpu_count() or 1)
            line["cpu_baseline"] = info
            line["other_configs"] = {"config2_dense_1x1M_G1024": config2_dense(ctx, torch, args)}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
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
    ap.add_argument("--targets-per-gpu", type=int, default=125000)
    ap.add_argument("--topk", type=int, default=100)
    ap.add_argument("--sweep-warps", type=int, default=16)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
