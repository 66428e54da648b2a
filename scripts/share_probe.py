"""Developer probe: the full config 3 (1,000,000 targets x G = 4096, top-100) at the query counts one GPU of an N-GPU
query-partitioned job sees (10,000 / N): total device time and the phases that do not shrink with the query count.
python scripts/share_probe.py [nq ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from prophylo_b200 import capi, synth

G, nt, k = 4096, 1_000_000, 100
nqs = [int(a) for a in sys.argv[1:]] or [1250, 10000]
P, Y, p = synth.make_queries(G, 10000, synth.SEED_BASE + 3, "mixed")
T = np.concatenate([synth.make_targets(G, nt // 8, synth.SEED_BASE + 103 + b, plant_from=Y, plant_frac=0.01) for b in range(8)])
c = capi.Context(0)
c.set_targets_bits(T, G)
flt = capi.make_filter(capi.FILTER_TOPK, k=k)
for nq, ov in [(n, o) for n in nqs for o in (0, 1)]:
    sel = np.arange(nq) * (10000 // nq)
    c.set_queries(P[sel], Y[sel], p[sel])
    try:
        c.set_option("pf_overlap", ov)
    except Exception:
        if ov:
            continue
    best = None
    for _ in range(4):
        c.run(flt)
        s = c.stats()
        if best is None or s["total_device_ms"] < best[0]:
            best = (s["total_device_ms"], {x: c.counter(x) for x in ("us_prefilter", "us_sweep", "us_exact", "us_staircase", "us_merge", "survivors")}, s["kernel_launches"])
    print(f"nq={nq} pf_overlap={ov}: {best[0]:.2f} ms = {nq * nt / best[0] * 1e3:.3e} pairs/s, launches {best[2]}, {best[1]}", flush=True)
c.close()
