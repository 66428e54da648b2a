"""BASELINE config 2 (1 query x 1,000,000 targets x 1024 genomes, dense): device time of the exact-table sweep."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from prophylo_b200 import capi, synth
G, nt = 1024, 1_000_000
P, Y, p = synth.make_queries(G, 1, synth.SEED_BASE + 2, "single")
T = synth.make_targets(G, nt, synth.SEED_BASE + 102, plant_from=Y, plant_frac=0.01)
c = capi.Context(0)
c.set_targets_bits(T, G); c.set_queries(P, Y, p)
ms = []
for _ in range(6):
    c.run(None); ms.append(c.stats()["total_device_ms"])
print("used_table", c.counter("used_table"), "ms", min(ms), "pairs/s", nt / min(ms) * 1e3, "GB/s alg", nt * (G // 8 + 24) / min(ms) / 1e6)
