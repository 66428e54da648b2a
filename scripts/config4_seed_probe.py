import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np
from prophylo_b200 import capi, synth
G, nq, nt = 2048, 2000, 500000
P, Y, p = synth.make_queries(G, nq, synth.SEED_BASE + 4, "prefix")
T = synth.make_targets(G, nt, synth.SEED_BASE + 104, plant_from=Y, plant_frac=0.01)
c = capi.Context(0); c.set_option("sweep_warps", 16)
c.set_targets_bits(T, G); c.set_queries(P, Y, p)
flt = capi.make_filter(capi.FILTER_TOPK, k=2)
ref = None
for sm, c1 in [(5, 24), (128, 512), (512, 512), (512, 2048), (1024, 4096), (2048, 4096)]:
    c.set_option("seed_mult", sm); c.set_option("chunk1_mult", c1)
    best = 1e9
    for _ in range(2):
        c.run(flt); best = min(best, c.stats()["total_device_ms"])
    h = c.fetch(); chk = (int(h["t"].astype(np.uint64).sum()), int(h["depth"].astype(np.uint64).sum())); ref = ref or chk
    print(f"seed_mult {sm} chunk1_mult {c1}: {best:.2f} ms prefilter {c.counter('us_prefilter')/1e3:.2f} surv {c.counter('survivors')} same {chk == ref}", flush=True)
