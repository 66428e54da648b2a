"""One top-k run with the tensor-core chunk screen, for ncu.  python scripts/tc_prof.py <prefilter_tc> <tc_exact> [nq] [nt]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prophylo_b200 import capi, synth

tc, exact = int(sys.argv[1]), int(sys.argv[2])
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
nt = int(sys.argv[4]) if len(sys.argv) > 4 else 30000
G, k = 4096, 100
P, Y, p = synth.make_queries(G, nq, synth.SEED_BASE + 3, "mixed")
T = synth.make_targets(G, nt, synth.SEED_BASE + 103, plant_from=Y, plant_frac=0.01)
c = capi.Context(0)
c.set_option("sweep_warps", 16)
c.set_option("prefilter_tc", tc)
c.set_option("tc_exact", exact)
c.set_targets_bits(T, G)
c.set_queries(P, Y, p)
c.run(capi.make_filter(capi.FILTER_TOPK, k=k))
print(c.stats()["total_device_ms"], c.counter("us_prefilter"), c.counter("survivors"), c.counter("tc_launches"))
