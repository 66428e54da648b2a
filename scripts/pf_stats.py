"""Developer probe (PF_STATS build): per-chunk screen pass rates of the pre-filter on the config-3 workload."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from prophylo_b200 import capi, synth
G, nq, nt = 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 2048, int(sys.argv[2]) if len(sys.argv) > 2 else 125000
P, Y, p = synth.make_queries(G, nq, synth.SEED_BASE + 3, "mixed")
T = synth.make_targets(G, nt, synth.SEED_BASE + 103, plant_from=Y, plant_frac=0.01)
c = capi.Context(0)
c.set_option("sweep_warps", 16)
c.set_targets_bits(T, G)
c.set_queries(P, Y, p)
lib = capi.load()
out = (ctypes.c_ulonglong * (8 * 64))()
c.run(capi.make_filter(capi.FILTER_TOPK, k=100))
lib.ppp_debug_pfstats(out, 1)
a = np.array(out[:], dtype=np.float64).reshape(8, 64)
tot = a[0][:32]
print("surv", c.counter("survivors"), "full", c.counter("full_pairs"), "pairs", nq * nt)
print("chunk  units     pass128(cy)  pass128(any)  pass256     pass512     final")
for ch in range(32):
    print(f"{ch:3d} {tot[ch]:12.0f} {a[1][ch]/tot[ch]:10.5f} {a[2][ch]/tot[ch]:10.5f} {a[3][ch]/tot[ch]:10.5f} {a[4][ch]/tot[ch]:10.5f} {a[5][ch]/tot[ch]:10.6f}")
print("sum per pair: pass128", a[1].sum() / tot[0], "pass128any", a[2].sum() / tot[0], "pass256", a[3].sum() / tot[0], "pass512", a[4].sum() / tot[0], "final", a[5].sum() / tot[0])
