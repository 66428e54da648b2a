"""Developer probe: time the sweep kernel in ALL mode on synthetic shapes (CUDA-event time from the library)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from prophylo_b200 import capi, synth

def run(G, kind, nq, nt, warps=8, reps=3, flt=None):
    P, Y, p = synth.make_queries(G, nq, 1, kind)
    T = synth.make_targets(G, nt, 2, plant_from=Y, plant_frac=0.01)
    c = capi.Context(0)
    c.set_option("sweep_warps", warps)
    c.set_targets_bits(T, G)
    c.set_queries(P, Y, p)
    best = None
    for r in range(reps):
        c.run(flt)
        s = c.stats()
        if best is None or s["sweep_kernel_ms"] < best["sweep_kernel_ms"]:
            best = s
    pairs = nq * nt
    print(f"G={G} {kind} nq={nq} nt={nt} warps={warps}: sweep {best['sweep_kernel_ms']:.2f} ms total {best['total_device_ms']:.2f} ms "
          f"-> {pairs / best['sweep_kernel_ms'] / 1e6:.2f} Mpairs/s/ms... = {pairs / (best['sweep_kernel_ms'] * 1e-3):.3e} pairs/s; "
          f"cand/pair {best['candidates'] / pairs:.1f} acc/pair {best['accurate_evals'] / pairs:.2f} exact {best['exact_pairs']} "
          f"surv {c.counter('survivors')} full {c.counter('full_pairs')} launches {best['kernel_launches']} us: sweep {c.counter('us_sweep')} exact {c.counter('us_exact')} stair {c.counter('us_staircase')} merge {c.counter('us_merge')}", flush=True)
    c.close()

if __name__ == "__main__":
    tk = capi.make_filter(capi.FILTER_TOPK, k=100)
    run(4096, "mixed", 2048, 65536, warps=16, flt=tk, reps=2)
    run(1024, "single", 1, 1000000, warps=16, flt=tk)
