"""Throughput of the native `.bla` reader (ppp_read_bla / ppp_read_bla_cached, prophylo_b200/csrc/ppp_bla.cpp) on synthetic
BLAST caches of the reference's format (query_gi, subject_gi, e-value, subject_taxon; util/cache_blast_result.pl:86).  Host
code only -- no GPU needed.  python scripts/bla_reader_bench.py [n_files] [hits_per_file] [genomes]"""
import os, sys, tempfile, time, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from prophylo_b200 import capi

n_files = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
hits = int(sys.argv[2]) if len(sys.argv) > 2 else 500
G = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
rng = np.random.default_rng(1)
universe = [str(t) for t in np.sort(rng.choice(np.arange(100, 2_000_000), size=G, replace=False))]
foreign = [str(t) for t in range(3_000_000, 3_000_400)]
d = tempfile.mkdtemp(prefix="bla_bench_")
try:
    t0 = time.perf_counter()
    paths, nbytes = [], 0
    uni = np.array(universe + foreign)
    for i in range(n_files):
        m = int(rng.integers(hits // 2, hits * 3 // 2))
        tax = uni[rng.integers(0, uni.size, size=m)]          # repeated taxa (paralogues) and foreign taxa as in real caches
        subj = rng.integers(1, 10**9, size=m)
        ev = rng.random(m) * 1e-5
        txt = "".join(f"{1000 + i}\t{s}\t{e:.2e}\t{t}\n" for s, e, t in zip(subj, ev, tax))
        p = os.path.join(d, f"{1000 + i}.bla")
        with open(p, "w") as f:
            f.write(txt)
        paths.append(p)
        nbytes += len(txt)
    print(f"wrote {n_files} files, {nbytes / 1e6:.1f} MB, {time.perf_counter() - t0:.1f} s", flush=True)
    ncpu = os.cpu_count() or 1
    for nth in sorted({1, min(8, ncpu), ncpu}):
        t0 = time.perf_counter()
        idx, raw, off = capi.read_bla_files(paths, universe, nthreads=nth)
        dt = time.perf_counter() - t0
        print(f"ppp_read_bla, {nth:3d} threads: {dt * 1e3:8.1f} ms = {n_files / dt:9.0f} files/s, {nbytes / dt / 1e6:7.1f} MB/s, "
              f"{idx.size} list entries ({idx.size / dt / 1e6:.1f} M entries/s)", flush=True)
    cache = os.path.join(d, "packed.cache")
    info = {}
    t0 = time.perf_counter()
    a = capi.read_bla_files(paths, universe, nthreads=ncpu, cache=cache, info=info)
    t_cold = time.perf_counter() - t0
    assert not info["from_cache"]
    t0 = time.perf_counter()
    b = capi.read_bla_files(paths, universe, nthreads=ncpu, cache=cache, info=info)
    t_warm = time.perf_counter() - t0
    assert info["from_cache"] and all(np.array_equal(x, y) for x, y in zip(a, b))
    print(f"ppp_read_bla_cached: first call (parse + write cache, {os.path.getsize(cache) / 1e6:.1f} MB) {t_cold * 1e3:.1f} ms; "
          f"second call (cache validated against the files' sizes and mtimes, then read) {t_warm * 1e3:.1f} ms = {n_files / t_warm:.0f} files/s", flush=True)
finally:
    shutil.rmtree(d, ignore_errors=True)
