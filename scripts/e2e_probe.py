"""Where the end-to-end step's time goes beyond the resident run: times each call of the e2e sequence separately (host clock,
synchronised) at config 3 with nt targets.  Developer aid; not part of the bench."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from prophylo_b200 import capi, synth

nq, nt, G, k = int(sys.argv[1]), int(sys.argv[2]), 4096, 100
P, Y, p = synth.make_queries(G, nq, synth.SEED_BASE + 3, "mixed")
T = synth.make_targets(G, nt, synth.SEED_BASE + 103, plant_from=Y, plant_frac=0.01)
Tp = torch.from_numpy(T).pin_memory().numpy()
Pp, Yp = torch.from_numpy(P).pin_memory().numpy(), torch.from_numpy(Y).pin_memory().numpy()
out = torch.empty(nq * k * capi.HIT_DTYPE.itemsize, dtype=torch.uint8).pin_memory().numpy().view(capi.HIT_DTYPE)
ctx = capi.Context(0)
flt = capi.make_filter(capi.FILTER_TOPK, k=k)


def clock(f):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = f()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0), r


for streamed in (False, True, False, True):
    ctx.set_targets_bits(Tp, G)
    ctx.score(Pp, Yp, p, flt, out=out)
    a, _ = clock(lambda: ctx.set_targets_bits(Tp, G, streamed=streamed))
    b, _ = clock(lambda: ctx.set_queries(Pp, Yp, p))
    c, _ = clock(lambda: ctx.run(flt))
    d, _ = clock(lambda: ctx.fetch(out=out))
    e, _ = clock(lambda: (ctx.set_targets_bits(Tp, G, streamed=streamed), ctx.score(Pp, Yp, p, flt, out=out)))
    f, _ = clock(lambda: ctx.run(flt))
    print(f"streamed={streamed}: set_targets {a:.2f} ms, set_queries {b:.2f}, run {c:.2f} (device {ctx.stats()['total_device_ms']:.2f}), fetch {d:.2f}; "
          f"whole e2e sequence {e:.2f}; resident run again {f:.2f}")
