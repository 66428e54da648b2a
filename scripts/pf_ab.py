"""Developer probe: config-3 shaped top-k run (argv[1] is ignored: there is one pre-filter); prints phase times.
With a -DPF2_STATS build (PPP_B200_LIB=ab/...) also the queue statistics of the version-2 pre-filter."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from prophylo_b200 import capi, synth
ver = int(sys.argv[1]) if len(sys.argv) > 1 else 2
G = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
nt = int(sys.argv[4]) if len(sys.argv) > 4 else 125000
kind = sys.argv[5] if len(sys.argv) > 5 else "mixed"
k = int(sys.argv[6]) if len(sys.argv) > 6 else 100
seed = {"prefix": 4, "family": 5, "single": 2}.get(kind, 3)
P, Y, p = synth.make_queries(G, nq, synth.SEED_BASE + seed, kind)
T = synth.make_targets(G, nt, synth.SEED_BASE + 100 + seed, plant_from=Y, plant_frac=0.01)
c = capi.Context(0)
c.set_option("sweep_warps", 16)
c.set_targets_bits(T, G)
c.set_queries(P, Y, p)
lib = capi.load()
out = (ctypes.c_ulonglong * 8)()
have_stats = hasattr(lib, "ppp_debug_pf2stats")
if have_stats:
    lib.ppp_debug_pf2stats.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
    lib.ppp_debug_pf2stats(c._h, out, 1)
flt = capi.make_filter(capi.FILTER_TOPK, k=k)
res = None
for r in range(3):
    c.run(flt)
    s = c.stats()
    print(f"v{ver} G={G} {nq}x{nt}: total {s['total_device_ms']:.2f} ms sweep {s['sweep_kernel_ms']:.2f} prefilter {c.counter('us_prefilter')/1e3:.2f} "
          f"exact {c.counter('us_exact')/1e3:.2f} stair {c.counter('us_staircase')/1e3:.2f} merge {c.counter('us_merge')/1e3:.2f} "
          f"surv {c.counter('survivors')} full {c.counter('full_pairs')} -> {nq*nt/s['total_device_ms']/1e6:.3f} Gpairs/s", flush=True)
    if have_stats and r == 0:
        lib.ppp_debug_pf2stats(c._h, out, 1)
        a = list(out)
        print(f"   pf2 stats: pushes/pair {a[1]/(nq*nt):.4f} drains {a[2]} drain-iters {a[3]} word-passes/pair {a[4]/(nq*nt):.5f}")
hits = c.fetch()
print("checksum", len(hits), int(hits["t"].astype(np.uint64).sum()), int(hits["depth"].astype(np.uint64).sum()), float(np.log(np.maximum(hits["score"], 1e-300)).sum()))
c.close()
