"""Throughput of the native profile-family reader + bit-plane packer (ppp_read_profiles, prophylo_b200/csrc/ppp_pack.cpp) on
synthetic profile files of the reference's format (taxon <tab> 0/1 with a header line; lib/PhyloProf/Profile.pm:433-514).
Host code only -- no GPU needed.  python scripts/profile_reader_bench.py [n_files] [genomes]"""
import os, sys, tempfile, time, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from prophylo_b200 import capi

n_files = int(sys.argv[1]) if len(sys.argv) > 1 else 4500
G = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
rng = np.random.default_rng(2)
taxa = np.sort(rng.choice(np.arange(100, 2_000_000), size=G, replace=False))
d = tempfile.mkdtemp(prefix="prof_bench_")
try:
    paths, nbytes = [], 0
    for i in range(n_files):
        dens = float(np.exp(rng.uniform(np.log(0.002), np.log(0.2))))
        y = rng.random(G) < dens
        txt = "taxon\tvalue\n" + "".join(f"{t}\t{int(v)}\n" for t, v in zip(taxa, y))
        p = os.path.join(d, f"fam{i}.profile")
        with open(p, "w") as f:
            f.write(txt)
        paths.append(p)
        nbytes += len(txt)
    print(f"wrote {n_files} profile files of {G} taxa, {nbytes / 1e6:.1f} MB", flush=True)
    ncpu = os.cpu_count() or 1
    for nth in sorted({1, ncpu}):
        t0 = time.perf_counter()
        universe, P, Y, p, yes, total = capi.read_profile_files(paths, header=True, nthreads=nth)
        dt = time.perf_counter() - t0
        print(f"ppp_read_profiles, {nth:3d} threads: {dt * 1e3:8.1f} ms = {n_files / dt:8.0f} files/s, {nbytes / dt / 1e6:7.1f} MB/s "
              f"(universe {len(universe)}, planes {P.shape}, ctypes copy-out included)", flush=True)
finally:
    shutil.rmtree(d, ignore_errors=True)
