"""Developer probe: top-k schedule parameters on the bench workload (10k x 125k x G=4096, k=100)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from prophylo_b200 import capi, synth
G, nq, nt = 4096, 10000, 125000
NB, NQ = int(os.environ.get("NBLOCKS", "1")), int(os.environ.get("NQ", "10000"))  # bench.py's blocks of 125k targets; queries used
P, Y, p = synth.make_queries(G, nq, synth.SEED_BASE + 3, "mixed")
T = np.concatenate([synth.make_targets(G, nt, synth.SEED_BASE + 103 + 1000 * b, plant_from=Y, plant_frac=0.01) for b in range(NB)])
P, Y, p = P[:NQ], Y[:NQ], p[:NQ]
c = capi.Context(0)
c.set_option("sweep_warps", 16)
if os.environ.get("LPL"):
    c.set_option("launch_pairs_log2", int(os.environ["LPL"]))
c.set_targets_bits(T, G)
c.set_queries(P, Y, p)
flt = capi.make_filter(capi.FILTER_TOPK, k=100)
ref = None
import itertools
respecs = [int(x) for x in os.environ.get("RESPEC", "8").split(",")]
combos = [(9, 0, 24, 5000), (9, 4, 24, 5000), (9, 2, 24, 5000), (5, 4, 24, 5000), (5, 4, 48, 5000), (9, 4, 48, 5000), (9, 4, 96, 5000),
          (5, 2, 48, 5000), (3, 4, 24, 5000), (5, 4, 48, 20000), (5, 4, 48, 1000), (7, 3, 36, 5000)]
if len(sys.argv) > 1:
    combos = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]]
for sm, c0, c1, eps, rs in [(a, b, cc, d, e) for (a, b, cc, d) in combos for e in respecs]:
    c.set_option("respec_mult", rs)
    c.set_option("seed_mult", sm); c.set_option("chunk0_mult", c0); c.set_option("chunk1_mult", c1); c.set_option("spec_eps_ppm", eps)
    best = None
    for _ in range(2):
        c.run(flt)
        s = c.stats()
        if best is None or s["total_device_ms"] < best[0]:
            best = (s["total_device_ms"], s["sweep_kernel_ms"], c.counter("us_prefilter") / 1e3, c.counter("us_exact") / 1e3,
                    c.counter("us_staircase") / 1e3, c.counter("us_merge") / 1e3, c.counter("survivors"), c.counter("spec_rank"), c.counter("spec_failed"))
    h = c.fetch()
    chk = (int(h["t"].astype(np.uint64).sum()), int(h["depth"].astype(np.uint64).sum()))
    ref = ref or chk
    print(f"respec {rs} seed {sm}k chunk0 {c0}k chunk1 {c1}k eps {eps}ppm: total {best[0]:.2f} ms sweep {best[1]:.2f} prefilter {best[2]:.2f} exact {best[3]:.2f} stair {best[4]:.2f} "
          f"merge {best[5]:.2f} surv {best[6]} rank {best[7]} failed {best[8]} same_result {chk == ref}", flush=True)
