"""A/B probe: dense (ALL) and top-k timings for the library selected by PPP_B200_LIB."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prophylo_b200 import capi, synth
def run(G, kind, nq, nt, flt, tag, table=-1):
    P, Y, p = synth.make_queries(G, nq, 1, kind); T = synth.make_targets(G, nt, 2, plant_from=Y, plant_frac=0.01)
    c = capi.Context(0); c.set_option("sweep_warps", 16); c.set_option("table_mode", table); c.set_targets_bits(T, G); c.set_queries(P, Y, p)
    best = None
    for _ in range(3):
        c.run(flt); s = c.stats()
        if best is None or s["total_device_ms"] < best["total_device_ms"]: best = s
    print(f"{tag}: total {best['total_device_ms']:.2f} ms sweep {best['sweep_kernel_ms']:.2f} ms -> {nq*nt/best['total_device_ms']/1e6:.3f} Gpairs/s acc/pair {best['accurate_evals']/(nq*nt):.3f} table {c.counter('used_table')} precomp_us {c.counter('us_staircase')}", flush=True)
    c.close()
run(4096, "mixed", 512, 8192, None, "dense G4096 512x8192")
run(1024, "single", 1, 1000000, None, "dense G1024 1x1M general", table=0)
run(1024, "single", 1, 1000000, None, "dense G1024 1x1M table", table=1)
run(4096, "mixed", 16, 125000, None, "dense G4096 16x125k general", table=0)
run(4096, "mixed", 16, 125000, None, "dense G4096 16x125k table", table=1)
run(4096, "mixed", 2048, 65536, capi.make_filter(capi.FILTER_TOPK, k=100), "topk G4096 2048x65536")
