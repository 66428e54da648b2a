"""BASELINE.json configs 2, 4 and 5 on one GPU (SURVEY.md section 8d shapes; config 3 is bench.py's headline workload).
Prints one JSON object; device time from the library's CUDA events, inputs resident.  python scripts/other_configs.py [scale]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from prophylo_b200 import capi, synth

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
out = {}
c = capi.Context(0)
c.set_option("sweep_warps", 16)


def timed(flt, reps=3):
    best = None
    for _ in range(reps):
        c.run(flt)
        s = c.stats()
        if best is None or s["total_device_ms"] < best["total_device_ms"]:
            best = s
    return best


# config 2: 1 query x 1M targets x 1024 genomes, every pair reported (dense) and top-100
G, nq, nt = 1024, 1, int(1_000_000 * scale)
P, Y, p = synth.make_queries(G, nq, synth.SEED_BASE + 2, "single")
T = synth.make_targets(G, nt, synth.SEED_BASE + 102, plant_from=Y, plant_frac=0.01)
c.set_targets_bits(T, G); c.set_queries(P, Y, p)
s = timed(None)
out["config2_dense"] = {"shape": f"{nq} x {nt} x G={G}", "ms": s["total_device_ms"], "pairs_per_s": nq * nt / s["total_device_ms"] * 1e3,
                        "mode": "exact table" if c.counter("used_table") else "general sweep",
                        "hbm_GBps_algorithmic": (nt * G / 8 + nt * 32) / s["total_device_ms"] / 1e6}
s = timed(capi.make_filter(capi.FILTER_TOPK, k=100))
out["config2_top100"] = {"ms": s["total_device_ms"], "pairs_per_s": nq * nt / s["total_device_ms"] * 1e3}

# config 4: ppp2d shape, 2k nested prefix queries x 500k targets x 2048 genomes, best non-self target per query (k = 2)
G, nq, nt = 2048, 2000, int(500_000 * scale)
P, Y, p = synth.make_queries(G, nq, synth.SEED_BASE + 4, "prefix")
T = synth.make_targets(G, nt, synth.SEED_BASE + 104, plant_from=Y, plant_frac=0.01)
c.set_targets_bits(T, G); c.set_queries(P, Y, p)
s = timed(capi.make_filter(capi.FILTER_TOPK, k=2))
out["config4_ppp2d_top2"] = {"shape": f"{nq} x {nt} x G={G}", "ms": s["total_device_ms"], "pairs_per_s": nq * nt / s["total_device_ms"] * 1e3,
                             "prefilter_ms": c.counter("us_prefilter") / 1e3, "survivors": c.counter("survivors"),
                             "spec_rank": c.counter("spec_rank"), "spec_failed": c.counter("spec_failed")}

# config 5: TIGRFAM-scale all-pairs, 4.5k x 4.5k x 8192 genomes (same set on both sides: the diagonal underflows to 0),
# ppp_cutoff threshold filter (printed %.3f log10 score < 0) and top-100
G, nq = 8192, int(4500 * min(1.0, scale * 4))
P, Y, p = synth.make_queries(G, nq, synth.SEED_BASE + 5, "family")
c.set_targets_bits(Y, G); c.set_queries(P, Y, p)
s = timed(capi.make_filter(capi.FILTER_THRESH_LOG10, thresh=-0.0005), reps=2)
out["config5_thresh"] = {"shape": f"{nq} x {nq} x G={G}", "ms": s["total_device_ms"], "pairs_per_s": nq * nq / s["total_device_ms"] * 1e3,
                         "hits": int(c.result_count()) if hasattr(c, "result_count") else None, "exact_pairs": s["exact_pairs"],
                         "prefilter_ms": c.counter("us_prefilter") / 1e3, "sweep_ms": c.counter("us_sweep") / 1e3,
                         "exact_ms": c.counter("us_exact") / 1e3, "rank_ms": c.counter("us_merge") / 1e3,
                         "survivors": c.counter("survivors")}
s = timed(capi.make_filter(capi.FILTER_TOPK, k=100), reps=2)
out["config5_top100"] = {"ms": s["total_device_ms"], "pairs_per_s": nq * nq / s["total_device_ms"] * 1e3,
                         "prefilter_ms": c.counter("us_prefilter") / 1e3, "survivors": c.counter("survivors")}
print(json.dumps(out, indent=1))
