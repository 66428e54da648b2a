"""Developer probe: where the end-to-end (host buffers) step of the bench workload spends its time."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from prophylo_b200 import capi, synth
G, nq, nt, k = 4096, 10000, 125000, 100
P, Y, p = synth.make_queries(G, nq, synth.SEED_BASE + 3, "mixed")
T = synth.make_targets(G, nt, synth.SEED_BASE + 103, plant_from=Y, plant_frac=0.01)
pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
P, Y, p, T = pin(P), pin(Y), pin(p), pin(T)
out = torch.empty(nq * k * capi.HIT_DTYPE.itemsize, dtype=torch.uint8).pin_memory().numpy().view(capi.HIT_DTYPE)
c = capi.Context(0); c.set_option("sweep_warps", 16)
flt = capi.make_filter(capi.FILTER_TOPK, k=k)
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    c.set_targets_bits(T, G); torch.cuda.synchronize(); t1 = time.perf_counter()
    c.set_queries(P, Y, p); torch.cuda.synchronize(); t2 = time.perf_counter()
    c.run(flt); torch.cuda.synchronize(); t3 = time.perf_counter()
    c.fetch(out); torch.cuda.synchronize(); t4 = time.perf_counter()
    print(f"set_targets {1e3*(t1-t0):.2f} ms  set_queries {1e3*(t2-t1):.2f} ms  run {1e3*(t3-t2):.2f} ms (device {c.stats()['total_device_ms']:.2f})  fetch {1e3*(t4-t3):.2f} ms  total {1e3*(t4-t0):.2f}")
