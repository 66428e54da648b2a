"""One small sweep for ncu: python scripts/prof_case.py G kind nq nt warps mode(all|topk)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prophylo_b200 import capi, synth
G, kind, nq, nt, warps = int(sys.argv[1]), sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
mode = sys.argv[6] if len(sys.argv) > 6 else "all"
P, Y, p = synth.make_queries(G, nq, 1, kind)
T = synth.make_targets(G, nt, 2, plant_from=Y, plant_frac=0.01)
c = capi.Context(0)
c.set_option("sweep_warps", warps)
c.set_targets_bits(T, G)
c.set_queries(P, Y, p)
flt = capi.make_filter(capi.FILTER_TOPK, k=100) if mode == "topk" else None
for _ in range(2):
    c.run(flt)
print(c.stats())
