"""Tensor-core chunk screen (prefilter_tc) against the POPC screen: identical results, survivor / full-evaluation counts,
pre-filter time.  python scripts/tc_probe.py [nq] [nt] [k] [G]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from prophylo_b200 import capi, synth

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 512
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
k = int(sys.argv[3]) if len(sys.argv) > 3 else 20
G = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
kind = sys.argv[5] if len(sys.argv) > 5 else "mixed"
sort_block = int(sys.argv[6]) if len(sys.argv) > 6 else 0
P, Y, p = synth.make_queries(G, nq, synth.SEED_BASE + 3, kind)
T = synth.make_targets(G, nt, synth.SEED_BASE + 103, plant_from=Y, plant_frac=0.01)
if sort_block:  # targets ordered by popcount inside blocks (what a density-sorted layout would give the lookups)
    pc = synth.unpack_bits(T, G).sum(axis=1)
    order = np.concatenate([s0 + np.argsort(pc[s0:s0 + sort_block], kind="stable") for s0 in range(0, nt, sort_block)])
    T = np.ascontiguousarray(T[order])
c = capi.Context(0)
c.set_option("sweep_warps", 16)
c.set_targets_bits(T, G)
flt = capi.make_filter(capi.FILTER_TOPK, k=k)
out = {}
res = {}
for tc in (0, 1, 3, 7):
    c.set_option("prefilter_tc", tc & 3)
    c.set_option("tc_exact", 0 if tc == 7 else 1)
    c.set_queries(P, Y, p)
    best = None
    for rep in range(3):
        c.run(flt)
        s = c.stats()
        if best is None or s["total_device_ms"] < best[0]:
            best = (s["total_device_ms"], c.counter("us_prefilter"), c.counter("survivors"), c.counter("full_pairs"), c.counter("tc_launches"),
                    c.counter("us_sweep"))
    res[tc] = c.fetch().copy()
    out[f"tc{tc}"] = dict(total_ms=best[0], prefilter_ms=best[1] / 1e3, survivors=best[2], full_pairs=best[3], tc_launches=best[4], sweep_ms=best[5] / 1e3)
    print(json.dumps({f"tc{tc}": out[f"tc{tc}"]}), flush=True)
a, b = res[0], res[3]
same = a.size == b.size and all(np.array_equal(a[f], b[f]) for f in ("q", "t", "yes", "number", "depth"))
out["identical_lists"] = bool(same) and all(np.array_equal(res[0][f], res[1][f]) for f in ("q", "t", "yes", "number", "depth"))
out["identical_scores"] = bool(a.size == b.size and np.array_equal(a["score"], b["score"]))
print(json.dumps(out))
