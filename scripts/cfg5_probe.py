"""Developer probe: where config 5's top-100 run (4,500 x 4,500 x G = 8192) and the dense G = 4096 run spend their time.
PPP_B200_DEBUG_CHUNKS=1 python scripts/cfg5_probe.py [topk|dense|all]  -- per-chunk lines on stderr, phase counters on stdout."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prophylo_b200 import capi, synth

what = sys.argv[1] if len(sys.argv) > 1 else "all"
c = capi.Context(0)
KEYS = ("us_sweep", "us_prefilter", "us_exact", "us_staircase", "us_merge", "survivors", "full_pairs")


def timed(tag, flt, pairs, reps=3):
    best = None
    for _ in range(reps + 1):
        c.run(flt)
        s = c.stats()
        if best is None or s["total_device_ms"] < best[0]["total_device_ms"]:
            best = (s, {k: c.counter(k) for k in KEYS})
    s, k = best
    print(f"{tag}: {s['total_device_ms']:.3f} ms = {pairs / s['total_device_ms'] * 1e3:.3e} pairs/s; launches {s['kernel_launches']} "
          f"cand/pair {s['candidates'] / pairs:.1f} acc/pair {s['accurate_evals'] / pairs:.2f} exact {s['exact_pairs']} {k}", flush=True)


if what in ("topk", "all"):
    G, nq = 8192, 4500
    P, Y, p = synth.make_queries(G, nq, synth.SEED_BASE + 5, "family")
    c.set_targets_bits(Y, G)
    c.set_queries(P, Y, p)
    timed("config5 top-100", capi.make_filter(capi.FILTER_TOPK, k=100), nq * nq)
if what in ("dense", "all"):
    G, nq, nt = 4096, 1000, 20_000
    P, Y, p = synth.make_queries(G, nq, synth.SEED_BASE + 3, "mixed")
    T = synth.make_targets(G, nt, synth.SEED_BASE + 103, plant_from=Y, plant_frac=0.01)
    c.set_targets_bits(T, G)
    c.set_queries(P, Y, p)
    c.set_option("table_mode", 0)
    timed("dense G=4096 1000 x 20000", None, nq * nt)
c.close()
